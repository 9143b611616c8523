"""Run one d3fk op on the GPU through the C ABI and the same op through the CPU interpreter
(tests/op_interpreter.py), on identical inputs.  For bf16 ops the CPU side sees the bf16-rounded
inputs held in fp32."""
import torch

from denoising_diffusion_deep_fake_b200 import _lib
import op_interpreter as I

T_FIELDS = {"src0", "src1", "w", "out", "res", "dy", "x", "y", "act", "dx", "dres", "dst", "w_fwd", "w_dgrad", "bw_x", "bw_act"}
# fields that are dtype-typed only for some ops
_NOT_T = {(_lib.OP_PACK, "w"), (_lib.OP_NCHW2NHWC, "src"), (_lib.OP_CHANSUM, "out")}


def _is_t(kind, name):
    return name in T_FIELDS and (kind, name) not in _NOT_T


def run_both(kind, dtype, tensors, scalars, outputs, dev="cuda:0"):
    """tensors: name -> CPU tensor (fp32 for T-typed fields).  Returns {name: (gpu_result_fp32_cpu, cpu_result)}."""
    _lib.init(0)
    tdt = torch.float32 if dtype == _lib.F32 else torch.bfloat16
    gpu, cpu = {}, {}
    for n, t in tensors.items():
        if _is_t(kind, n):
            g = t.to(dev).to(tdt)
            gpu[n] = g
            cpu[n] = g.float().cpu().contiguous()
        else:
            gpu[n] = t.to(dev)
            cpu[n] = t.clone()
    fg = dict(scalars)
    fc = dict(scalars)
    fg["dtype"] = dtype
    fc["dtype"] = _lib.F32
    for n in tensors:
        fg[n] = gpu[n].data_ptr()
        fc[n] = cpu[n].data_ptr()
    if "dtype" not in {f[0] for f in _lib._PARAM_CLS[_lib._UNION_FIELD[kind]]._fields_}:
        fg.pop("dtype"), fc.pop("dtype")
    opg = _lib.make_op(kind, **fg)
    opc = _lib.make_op(kind, **fc)
    _lib.run_single(opg, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert _lib.load().d3fk_device_error_flag() == 0, "kernel watchdog tripped"
    I.run_ops([opc])
    return {n: (gpu[n].float().cpu(), cpu[n].float()) for n in outputs}


def rel_err(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()

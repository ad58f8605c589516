"""Developer tool: per-tile clock64 timeline of CTA 0 of gemm_tc_kernel (needs a library built with
`make -C isegprobe_b200/csrc EXTRA=-DISP_GEMM_TRACE`).  Prints, for the first tiles, when the MMA warp waited for / got the
accumulator, finished issuing, how long it waited for operands, and when epilogue warps 0 / 15 got the accumulator and
released it."""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import tc, _lib
dev, bf = "cuda", torch.bfloat16
M = 4 * 448 * 448
D, C, Dp, NQ = 404, 384, 416, 448
x = torch.randn(M, Dp, device=dev).to(bf); x[:, D:] = 0
st_in = torch.zeros(M, tc.stats_slots(D, bf, False), 2, device=dev); st_in[..., 1] = 50.0
Wq, gq, bq = [t.to(dev) for t in tc.pack_ln_linear(torch.randn(NQ, D) * 0.05, torch.zeros(NQ), torch.ones(D), torch.zeros(D))]
if os.environ.get("NOLN"):
    fn = lambda: tc.gemm(x, Wq, bias=bq, out_dtype=bf, N=NQ, K=D)
else:
    fn = lambda: tc.gemm(x, Wq, bias=bq, out_dtype=bf, N=NQ, K=D, ln_stats=st_in, ln_g=gq, ln_eps=1e-5)
L = _lib.lib()
buf = np.zeros((4, 64, 4), dtype=np.int64)
for _ in range(3):
    fn()
L.isp_gemm_trace_read(buf.ctypes.data_as(ctypes.c_void_p), 1)
fn()
L.isp_gemm_trace_read(buf.ctypes.data_as(ctypes.c_void_p), 1)
t0 = buf[0, 0, 0]
print("tile | mma: wait_acc got_acc issued opwait | epi0: wait got done | epi15: wait got done   (cycles from start)")
for t in range(40):
    m, e, f = buf[0, t], buf[1, t], buf[2, t]
    print(f"{t:3d} | {m[0]-t0:7d} {m[1]-t0:7d} {m[2]-t0:7d} {m[3]:6d} | {e[0]-t0:7d} {e[1]-t0:7d} {e[2]-t0:7d} | {f[0]-t0:7d} {f[1]-t0:7d} {f[2]-t0:7d}")

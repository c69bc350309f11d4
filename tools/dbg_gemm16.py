"""Bring-up aid for csrc/tc_gemm16.cu: the split-fp16 Linear against fp64 and against the 3xTF32 kernel, plus timings."""
import os, sys, importlib, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("recommender-baseline-model_b200")
ops = importlib.import_module("recommender-baseline-model_b200.ops")
L = importlib.import_module("recommender-baseline-model_b200.lib")
dev = torch.device("cuda:0")


def run(x, w, b, res, tok, act, pA, pB, g16):
    os.environ["RBM_LINEAR_GEMM16"] = "1" if g16 else "0"
    return ops.linear(x, w, b, residual=res, row_tok=tok, act=act, pA=pA, siteA=3, pB=pB, siteB=4, seed=77)


def check():
    torch.manual_seed(0)
    for (M, N, K) in [(512, 128, 128), (1000, 256, 256), (4099, 1024, 256), (3000, 256, 1024), (777, 384, 128), (20000, 128, 192)]:
        x = torch.randn(M, K, device=dev) * 3
        w = torch.randn(N, K, device=dev) * 0.05
        b = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        tok = (torch.rand(M, device=dev) > 0.2).long()
        ref = (x.double() @ w.double().t() + b.double())
        y = run(x, w, b, None, None, L.ACT_NONE, 0.0, 0.0, True)
        y0 = run(x, w, b, None, None, L.ACT_NONE, 0.0, 0.0, False)
        e = (y.double() - ref).abs().max().item() / ref.abs().max().item()
        e0 = (y0.double() - ref).abs().max().item() / ref.abs().max().item()
        # full epilogue: identical masks => the two paths must agree to rounding
        ya = run(x, w, b, res, tok, L.ACT_GELU_TANH, 0.1, 0.2, True)
        yb = run(x, w, b, res, tok, L.ACT_GELU_TANH, 0.1, 0.2, False)
        ea = (ya - yb).abs().max().item()
        # backward-data
        xg = x.clone().requires_grad_(True)
        os.environ["RBM_LINEAR_GEMM16"] = "1"
        dy = torch.randn(M, N, device=dev)
        ops.linear(xg, w, b).backward(dy)
        dref = dy.double() @ w.double()
        ed = (xg.grad.double() - dref).abs().max().item() / dref.abs().max().item()
        print(f"M={M} N={N} K={K}: fwd rel {e:.2e} (3xTF32 {e0:.2e}) | epilogue diff {ea:.2e} | dx rel {ed:.2e}")


def time_():
    for (M, N, K) in [(102400, 256, 256), (102400, 1024, 256), (102400, 256, 1024), (102400, 768, 256), (204800, 128, 128), (204800, 384, 128)]:
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * 0.05
        b = torch.randn(N, device=dev)
        for g16 in (True, False):
            os.environ["RBM_LINEAR_GEMM16"] = "1" if g16 else "0"
            for act in (L.ACT_NONE, L.ACT_GELU_TANH):
                with torch.no_grad():
                    for _ in range(3):
                        ops.linear(x, w, b, act=act, pA=0.1 if act else 0.0, siteA=3)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        ops.linear(x, w, b, act=act, pA=0.1 if act else 0.0, siteA=3)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"M={M} N={N} K={K} gemm16={int(g16)} act={act}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s  {(M*K+M*N)*4/ms/1e6:.0f} GB/s")


def prof():
    M, N, K = 102400, 1024, 256
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    os.environ["RBM_LINEAR_GEMM16"] = "1"
    with torch.no_grad():
        for _ in range(2):
            ops.linear(x, w, b)
    torch.cuda.synchronize()


if __name__ == "__main__":
    {"check": check, "time": time_, "prof": prof}[sys.argv[1]]()

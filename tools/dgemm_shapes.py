"""Which cuBLAS fp64 kernel / layout is fastest for the ChFSI block product C (d x d, symmetric) times Y (d x b)?
torch.matmul on fp64 tensors goes to the same cublasDgemm heuristics eig.cu uses; the layouts below reach the NN / TN / NT
forms and the transposed problem (b x d result). CUDA-event timing, 200 products each, L2-warm like the real filter loop."""
import torch, json, sys
d = 2000
dev = torch.device("cuda:0")
torch.manual_seed(0)
C = torch.randn(d, d, dtype=torch.float64, device=dev)
C = (C + C.T).contiguous()
out = []
def timeit(name, fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    return us
for b in (160, 128, 192, 80, 40, 20):
    Yt = torch.randn(b, d, dtype=torch.float64, device=dev)      # row-major b x d == column-major d x b (eig.cu's layout)
    Y = Yt.T.contiguous()                                         # row-major d x b
    Ct = C.T                                                      # a transposed view: cuBLAS gets OP_T
    o1 = torch.empty(b, d, dtype=torch.float64, device=dev); o2 = torch.empty(d, b, dtype=torch.float64, device=dev)
    res = {"b": b, "gflop": 2 * d * d * b / 1e9}
    res["Yt@C (eig.cu: NN, m=d n=b)"] = timeit("a", lambda: torch.matmul(Yt, C, out=o1))
    res["Yt@C.T (OP_T on C)"] = timeit("b", lambda: torch.matmul(Yt, Ct, out=o1))
    res["C@Y (m=b n=d)"] = timeit("c", lambda: torch.matmul(C, Y, out=o2))
    res["C.T@Y"] = timeit("d", lambda: torch.matmul(Ct, Y, out=o2))
    res["C@Yt.T"] = timeit("e", lambda: torch.matmul(C, Yt.T, out=o2))
    best = min(v for k, v in res.items() if k not in ("b", "gflop"))
    res["best_tflops"] = res["gflop"] / best / 1e3 * 1e3
    out.append(res); print(json.dumps(res), flush=True)
json.dump(out, open("gpurun_out/dgemm_shapes.json", "w"), indent=1)

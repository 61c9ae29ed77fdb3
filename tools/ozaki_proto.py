#!/usr/bin/env python
"""First prototype (kept for the record; the kernel in csrc/ozaki.cu uses 6 radix-256 digits, see oracle/ozaki_oracle.py):
FP64-accurate GEMM on the INT8 tensor cores (Ozaki scheme), torch only.

C = A @ B with A [M,K] sliced per ROW and B [K,N] sliced per COLUMN into S signed 7-bit slices
(error-free: each slice is an exact truncation of the scaled remainder); the S(S+1)/2 slice products
with p+q < S are exact INT32 GEMMs (torch._int_mm), recombined in FP64.
Reports accuracy vs a float64 matmul and the time of the INT8 products alone.
"""
import sys
import torch

torch.manual_seed(0)
dev = "cuda"


def slices_rows(a, S):
    """a [R, K] float64 -> (list of S int8 [R, K], exponent scale [R,1]) with a ~= scale * sum_p 2^{-7(p+1)} s_p."""
    amax = a.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax))            # |a| <= 2^e
    scale = torch.exp2(e)
    r = a / scale                               # |r| <= 1 (exact: power-of-two scaling)
    out = []
    for _ in range(S):
        r = r * 128.0
        s = torch.trunc(r)
        s = s.clamp_(-127, 127)                 # r == +-128 only if |a| == 2^e exactly
        r = r - s
        out.append(s.to(torch.int8))
    return out, scale


def ozaki_matmul(a, b, S, time_it=False):
    sa, ea = slices_rows(a, S)
    sbt, eb = slices_rows(b.t().contiguous(), S)          # slice columns of B
    sb = [s.t().contiguous() for s in sbt]
    M, N = a.shape[0], b.shape[1]
    acc = torch.zeros(M, N, dtype=torch.float64, device=a.device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    int_ms = 0.0
    for d in range(S):                                   # diagonal p + q = d
        part = torch.zeros(M, N, dtype=torch.int32, device=a.device)
        ev0.record()
        for p in range(d + 1):
            part += torch._int_mm(sa[p], sb[d - p])
        ev1.record()
        torch.cuda.synchronize()
        int_ms += ev0.elapsed_time(ev1)
        acc += part.to(torch.float64) * (2.0 ** (-7 * (d + 2)))
    return acc * ea * eb.t(), int_ms


def main():
    M, K, N = 60000, 784, 256
    x = torch.rand(M, K, dtype=torch.float64, device=dev) * 2 - 1
    w = (torch.rand(K, N, dtype=torch.float64, device=dev) * 2 - 1) * 0.25
    ref = x @ w
    # K must be a multiple of 8 for _int_mm
    for S in (5, 6, 7, 8):
        c, ms = ozaki_matmul(x, w, S)
        err = ((c - ref).norm() / ref.norm()).item()
        emax = ((c - ref).abs().max() / ref.abs().max()).item()
        npairs = S * (S + 1) // 2
        print("fwd0  S=%d pairs=%2d  rel_fro=%.3e  max_abs/max=%.3e  int8 time %.3f ms (%.1f TOPS)" %
              (S, npairs, err, emax, ms, 2.0 * M * N * K * npairs / ms / 1e9), flush=True)
    # weight-gradient shape: C[784,256] = X^T[784,60000] . delta[60000,256], heavy-tailed delta
    delta = torch.randn(M, N, dtype=torch.float64, device=dev) * torch.exp(torch.randn(M, 1, dtype=torch.float64, device=dev) * 2) * 1e-5
    xt = x.t().contiguous()
    ref = xt @ delta
    for S in (6, 7, 8):
        # INT32 accumulation is exact only while K * 127^2 < 2^31  =>  K <= 133000; fine for K = 60000 per product,
        # but a diagonal sums up to S products -> accumulate each product in int32 separately (done in ozaki_matmul? no:
        # `part` sums d+1 products) -> split K into chunks of 16384 rows to stay exact
        acc = torch.zeros(K, N, dtype=torch.float64, device=dev)
        tot_ms = 0.0
        sa, ea = slices_rows(xt, S)
        sbt, eb = slices_rows(delta.t().contiguous(), S)
        sb = [s.t().contiguous() for s in sbt]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        CH = 15000
        for k0 in range(0, M, CH):
            for d in range(S):
                part = torch.zeros(K, N, dtype=torch.int32, device=dev)
                ev0.record()
                for p in range(d + 1):
                    part += torch._int_mm(sa[p][:, k0:k0 + CH].contiguous(), sb[d - p][k0:k0 + CH].contiguous())
                ev1.record(); torch.cuda.synchronize(); tot_ms += ev0.elapsed_time(ev1)
                acc += part.to(torch.float64) * (2.0 ** (-7 * (d + 2)))
        c = acc * ea * eb.t()
        err = ((c - ref).norm() / ref.norm()).item()
        print("dW0   S=%d rel_fro=%.3e  (int8 products incl. slicing copies %.2f ms)" % (S, err, tot_ms), flush=True)
    # reference: FP64 matmul time (cuBLAS DGEMM)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x @ w; ev0.record(); x @ w; ev1.record(); torch.cuda.synchronize()
    print("cuBLAS DGEMM fwd0: %.3f ms" % ev0.elapsed_time(ev1))


if __name__ == "__main__":
    main()

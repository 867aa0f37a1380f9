#!/usr/bin/env python
"""Measure the tensor-rate denominators MEASURED_PEAKS.json does not hold and write profiles/peaks_r2.json:

  int8_tops_burst     torch._int_mm 8192^3 (cuBLASLt int8, the library ceiling, best of 10 -- the int8 twin of the bf16 burst figure)
  int8_issue_tops     tcgen05.mma kind::i8 issue rate on 148 SMs from zero-filled shared memory (scratch/mma_rate.cu)
  fp4_issue_tops      tcgen05.mma kind::mxf4.block_scale issue rate, N = 240 (scratch/mxf4_probe.cu)
  bf16_issue_tflops   kind::f16 issue rate
  popc_gops           POPC throughput of the whole chip (uint32 popcounts per second, register-resident loop)
  own kernels         qt_gemm_i8 / qt_gemm_f4 at 8192 x 8192 x 8192 (CUDA-graph timed)

Run on the GPU box:  python profiles/measure_peaks.py  (needs nvcc for the two probes; same image as the build container)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def best_ms(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def probe(src, pattern):
    exe = "/tmp/" + os.path.basename(src).replace(".cu", "")
    r = subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", os.path.join(ROOT, src), "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0:
        return None, r.stderr[-400:]
    out = subprocess.run([exe], capture_output=True, text=True).stdout
    vals = {}
    for ln in out.splitlines():
        m = re.match(pattern, ln)
        if m:
            vals[(m.group(1).strip(), int(m.group(2)), int(m.group(3)))] = float(m.group(4))
    return vals, out


POPC_SRC = r"""
#include <cstdio>
#include <cstdint>
__global__ void k(uint32_t* out, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + blockIdx.x, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int i = 0; i < iters; ++i) { s0 += __popc(a ^ s3); s1 += __popc(a + s0); s2 += __popc(a ^ s1); s3 += __popc(a + s2); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  int iters = 100000;
  k<<<148 * 8, 1024>>>(d, 1000); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<<<148 * 8, 1024>>>(d, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("popc_gops %.1f\n", 4.0 * iters * 148 * 8 * 1024 / (ms * 1e-3) / 1e9);
  return 0;
}
"""


def main():
    dev = torch.device("cuda")
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    n = 8192
    a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    ms = best_ms(lambda: torch._int_mm(a, b))
    res["int8_tops_burst"] = round(2.0 * n ** 3 / ms / 1e9, 1)
    res["int8_how"] = "torch._int_mm 8192^3, best of 10, CUDA events"
    x = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    y = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    res["bf16_tflops_burst_here"] = round(2.0 * n ** 3 / best_ms(lambda: x @ y) / 1e9, 1)
    pat = r"^(\S[^N]*?)\s+N=\s*(\d+) blocks=\s*(\d+):.*?\(([\d.]+) Pop/s\)"
    v, raw = probe("scratch/mma_rate.cu", pat)
    if v:
        res["int8_issue_tops"] = round(v.get(("i8", 256, 148), 0) * 1e3, 1)
        res["fp8_issue_tops"] = round(v.get(("fp8 e4m3", 256, 148), 0) * 1e3, 1)
        res["bf16_issue_tflops"] = round(v.get(("bf16", 256, 148), 0) * 1e3, 1)
        res["tf32_issue_tflops"] = round(v.get(("tf32", 256, 148), 0) * 1e3, 1)
    res["mma_rate_raw"] = raw
    v, raw = probe("scratch/mxf4_probe.cu", pat)
    if v:
        res["fp4_issue_tops"] = round(v.get(("mxf4", 240, 148), 0) * 1e3, 1)
        res["fp4_tops_burst"] = res["fp4_issue_tops"]
    res["mxf4_probe_raw"] = raw
    with open("/tmp/popc.cu", "w") as f:
        f.write(POPC_SRC)
    r = subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "/tmp/popc.cu", "-o", "/tmp/popc"], capture_output=True, text=True)
    if r.returncode == 0:
        out = subprocess.run(["/tmp/popc"], capture_output=True, text=True).stdout
        m = re.search(r"popc_gops ([\d.]+)", out)
        if m:
            res["popc_gops"] = float(m.group(1))
            res["popc_how"] = "4 dependent-chain __popc per iteration per thread, 148 x 8 x 1024 threads"
    # own kernels at 8192^3
    import bench
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    out = torch.empty(n, n, device=dev)
    ai = torch.randint(-3, 4, (n, n), dtype=torch.int8, device=dev)
    wi = torch.randint(-3, 4, (n, n), dtype=torch.int8, device=dev)
    ms = bench.time_fn(torch, lambda: ops.gemm_i8(ai, True, n, wi, True, n, n, n, n, ops.make_epi(out, ldo=n), L.BACKEND_TCGEN05), iters=20, graph=True)
    res["qt_gemm_i8_8192_tops"] = round(2.0 * n ** 3 / ms / 1e9, 1)
    a4 = (torch.randint(0, 256, (n, n // 2), dtype=torch.uint8, device=dev) & 0xAA) | 0x22
    w4 = (torch.randint(0, 256, (n, n // 2), dtype=torch.uint8, device=dev) & 0xAA) | 0x22
    ms = bench.time_fn(torch, lambda: ops.gemm_f4(a4, n, w4, n, n, n, n, ops.make_epi(out, ldo=n)), iters=20, graph=True)
    res["qt_gemm_f4_8192_tops"] = round(2.0 * n ** 3 / ms / 1e9, 1)
    path = os.path.join(ROOT, "gpurun_out", "peaks_r2.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if not k.endswith("_raw")}, indent=1))


if __name__ == "__main__":
    main()

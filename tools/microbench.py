"""Write-bandwidth ceilings on B200 for the observation-row store pattern (run on the GPU box; not product code)."""
import ctypes as C
import json
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmicrobench.so")


def build():
    subprocess.check_call(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                           "-o", LIB, os.path.join(HERE, "microbench.cu")])


def timeit(fn, bytes_per_call, reps=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return bytes_per_call / (ms * 1e-3) / 1e9, ms * 1e3


def main():
    if not os.path.exists(LIB):
        build()
    L = C.CDLL(LIB)
    L.mb_tma_rows.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.mb_tma_block.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
    L.mb_lsu.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    n_rows, row = 65536, 1200
    nbuf = 4
    bufs = [torch.empty(n_rows * row, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    big2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    out = {}
    out["torch_fill_1GiB"] = timeit(lambda k: big.fill_(1), 1 << 30, 10)
    out["torch_copy_1GiB_rw"] = timeit(lambda k: big2.copy_(big), 2 << 30, 10)
    out["fill_78MB_ring4"] = timeit(lambda k: bufs[k % nbuf].fill_(1), n_rows * row)
    for streaming in (0, 1):
        for ctas in (148 * 4, 148 * 8, 148 * 16):
            out[f"lsu_st16_cs{streaming}_ctas{ctas}"] = timeit(
                lambda k: L.mb_lsu(bufs[k % nbuf].data_ptr(), n_rows * row, streaming, ctas, st), n_rows * row)
    for rpc in (64, 128, 256):
        out[f"tma_rows_1200B_rpc{rpc}"] = timeit(lambda k: L.mb_tma_rows(bufs[k % nbuf].data_ptr(), n_rows, row, rpc, 0, 0, st), n_rows * row)
        out[f"tma_rows_48+1152_rpc{rpc}"] = timeit(lambda k: L.mb_tma_rows(bufs[k % nbuf].data_ptr(), n_rows, row, rpc, 48, 0, st), n_rows * row)
        out[f"tma_rows_lsu48+1152_rpc{rpc}"] = timeit(lambda k: L.mb_tma_rows(bufs[k % nbuf].data_ptr(), n_rows, row, rpc, 48, 1, st), n_rows * row)
    for bpc in (19200, 38400, 76800, 153600):
        out[f"tma_block_{bpc}B"] = timeit(lambda k: L.mb_tma_block(bufs[k % nbuf].data_ptr(), n_rows * row, bpc, st), n_rows * row)
    for k, v in out.items():
        print(f"{k:40s} {v[0]:9.1f} GB/s  {v[1]:9.2f} us")
    json.dump(out, open(os.path.join(HERE, "..", "gpurun_out", "microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

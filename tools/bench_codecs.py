#!/usr/bin/env python3
"""Secondary benchmarks for BASELINE.json configs 3-5 (LZ4 blocks, BZip2 900 KB blocks, XZ/LZMA2 1 MiB streams) on ONE GPU.

bench.py stays the headline (config 2, Deflate).  Each workload prints one JSON line with the same keys: decompressed GB/s
with the batch resident in HBM, the HBM roofline fraction of the (single) kernel, and the CPU restatement's throughput.
Unit counts are scaled to one GPU / a few minutes (stated in `config`); distinct units are tiled on the device.
"""
import argparse
import ctypes as C
import bz2
import json
import lzma
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _lz4_unit(seed):
    import helpers as H
    rng = np.random.Generator(np.random.PCG64(seed))
    k = rng.random()
    raw = bytes(65536) if k < 0.1 else rng.integers(0, 256, 65536, dtype=np.uint8).tobytes() if k < 0.2 else H.textlike(65536, seed)
    return H.lz4_block_compress(raw), raw


def _bz2_unit(seed):
    import helpers as H
    raw = H.textlike(900000, seed)
    return bz2.compress(raw, 9), raw


def _xz_unit(seed):
    import helpers as H
    raw = H.textlike(1 << 20, seed)
    # the unit handed to the batched kernel is the raw LZMA2 stream of the block (XZ framing is parsed on the host)
    return lzma.compress(raw, format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA2, "preset": 6, "dict_size": 1 << 20}]), raw


WORKLOADS = {
    "lz4": dict(gen=_lz4_unit, seed0=3, codec="lz4_block", unit=65536, distinct=2048, units=262144, kernel="lz4_parse_kernel + lz4_exec_kernel",
                desc="LZ4 block mode: independent 64 KiB blocks (80 % text-like, 10 % zeros, 10 % incompressible), LZ4_compress_default"),
    "bzip2": dict(gen=_bz2_unit, seed0=4, codec="bzip2", unit=900000, distinct=64, units=2048, kernel="bzip2_kernel",
                  desc="BZip2: independent single-block 900 KB streams, bz2 level 9"),
    "xz": dict(gen=_xz_unit, seed0=5, codec="lzma2", unit=1 << 20, distinct=32, units=1184, kernel="lzma_kernel",
               desc="XZ/LZMA2: independent 1 MiB streams, preset 6, 1 MiB dictionary (raw LZMA2 payload of each XZ block)"),
}


def oracle_fn(name):
    import swco
    if name == "lz4":
        return lambda u: swco.lz4_block(u)
    if name == "bzip2":
        return lambda u: swco.bzip2_decompress(u)
    return lambda u: swco.lzma2_decompress_raw(u, 18)


def cpu_throughput(name, units, budget, threads):
    """pthreads inside oracle/batch_mt.c (no interpreter in the timed loop)"""
    import swco
    codec, aux = {"lz4": ("lz4_block", 0), "bzip2": ("bzip2", 0), "xz": ("lzma2", 18)}[name]
    sec, nbytes, fails = swco.batch_mt(codec, units, max(threads, 4), threads, aux)
    assert fails == 0
    total = max(int(budget / (sec / max(threads, 4))), threads)
    sec, nbytes, fails = swco.batch_mt(codec, units, total, threads, aux)
    assert fails == 0
    return nbytes / sec / 1e9, total, sec


def best_threads(name, units):
    ncpu = os.cpu_count() or 1
    sweep = {th: cpu_throughput(name, units, 1.0, th)[0] for th in sorted({max(1, ncpu >> k) for k in range(5)} | {min(ncpu, 16), min(ncpu, 24)})}
    return max(sweep, key=sweep.get), {str(k): round(v, 3) for k, v in sweep.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=list(WORKLOADS), required=True)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--units", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    W = WORKLOADS[args.workload]
    n_units = args.units or W["units"]
    distinct = min(W["distinct"], n_units)
    tile = max(n_units // distinct, 1)
    n_units = tile * distinct
    with Pool(min(os.cpu_count() or 1, 32)) as pool:
        pairs = pool.map(W["gen"], range(W["seed0"], W["seed0"] + distinct), chunksize=1)
    units = [p[0] for p in pairs]
    raws = [p[1] for p in pairs]

    import torch
    from swcompression_b200 import _lib
    from swcompression_b200.batch import Batch, pack_units
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    L = _lib.lib()
    buf, offs, lens = pack_units(units)
    stride = len(buf) - 64
    d_in = torch.cat([torch.from_numpy(buf[:stride]).to(dev).repeat(tile), torch.zeros(64, dtype=torch.uint8, device=dev)])
    all_off = (offs[None, :] + (np.arange(tile, dtype=np.uint64) * np.uint64(stride))[:, None]).reshape(-1)
    all_len = np.tile(lens, tile)
    aux = bytes([18] * n_units) if W["codec"] == "lzma2" else None
    b = Batch(W["codec"], np.zeros(1, dtype=np.uint8), all_off, all_len, W["unit"], device=str(dev), aux=aux)
    b.d_in = d_in
    total_in = int(lens.sum()) * tile
    total_out = n_units * W["unit"]
    for _ in range(args.warmup):
        b.run()
    st, ln, used = b.results()
    assert (st == 0).all() and (ln == W["unit"]).all(), (st[:8], ln[:8])
    fn = oracle_fn(args.workload)
    host = b.d_out[: min(distinct, 8) * ((W["unit"] + 15) // 16 * 16)].cpu().numpy()
    pu = (W["unit"] + 15) // 16 * 16
    for i in range(min(distinct, 8)):
        ost, oout, _ = fn(units[i])
        assert ost == 0 and oout == raws[i] and bytes(host[i * pu:i * pu + W["unit"]]) == oout, "parity vs oracle failed"
    launches0 = L.swc_kernel_launches()
    L.swc_timing_collect.argtypes = [C.c_void_p, C.c_int32]
    L.swc_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        b.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    tbuf = (C.c_float * 64)()
    nint = L.swc_timing_collect(tbuf, 64)
    L.swc_timing_enable(0)
    marks = [round(float(x), 3) for x in list(tbuf)[:nint]]
    launches = L.swc_kernel_launches() - launches0
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    achieved = (total_in + total_out) / (ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu:
        cores, sweep = best_threads(args.workload, units)
        v1, n1, d1 = cpu_throughput(args.workload, units, 4.0, 1)
        vN, nN, dN = cpu_throughput(args.workload, units, 8.0, cores)
        cpu = {"value": vN, "unit": "GB/s", "cores": cores, "logical_cpus": os.cpu_count(), "thread_sweep_GBps": sweep, "kind": "port",
               "single_thread_value": v1, "sample": f"{nN} units in {dN:.1f} s on {cores} pthreads (oracle/batch_mt.c)"}
    print(json.dumps({
        "metric": "decompressed_GB_per_s", "value": total_out / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": W["desc"], "units": n_units, "unit_bytes": W["unit"], "distinct_units": distinct,
                   "compressed_bytes": total_in, "decompressed_bytes": total_out},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": W["kernel"], "kernel_ms": ms, "algorithmic_bytes_per_launch": total_in + total_out,
                     "intervals_between_timing_marks_ms": marks},
        "cpu_baseline": cpu, "gpu_launches": int(launches)}))


if __name__ == "__main__":
    main()

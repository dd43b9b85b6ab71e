#!/usr/bin/env python
"""bench.py -- encoder input MB/s of the Brotli hot path on B200 (BASELINE.json metric).

A "step" = one pass of the hot path over one batch: at N GPUs every rank compresses its own
100 000 000-byte enwik8-shaped synthetic text stream (configs[1] of BASELINE.json: quality 5,
lgwin 22, one-shot, bit-exact to the reference).  Independent streams share nothing, so the
path shards across GPUs without a data-path collective (weak scaling); the compressed shards are
gathered to rank 0 over NCCL at the end of every step (variable-size gather).

  value : whole-job input MB/s with the input already resident in HBM (BrotliB200CompressDevice)
  e2e   : same metric through the reference-facing C ABI call BrotliEncoderCompress with pinned
          HOST buffers: H2D of the input and D2H of the compressed bytes inside the timed region
  --impl reference : the reference's own CPU encoder (oracle/_ref, built from /root/reference by
          oracle/Makefile; falls back to the oracle port) timed on the host cores for the same config
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

QUALITY, LGWIN = 5, 22
WORKLOAD_BYTES = 100_000_000
WORKLOAD = "100 MB enwik8-shaped synthetic text (tests/corpus.py synth_text, seed 20250922+rank), quality 5, lgwin 22, one stream per GPU"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_input(rank, nbytes=WORKLOAD_BYTES):
    from corpus import synth_text
    t = time.time()
    d = synth_text(nbytes, seed=20250922 + rank)
    log("rank %d: generated %d bytes in %.1fs" % (rank, len(d), time.time() - t))
    return d


def ref_lib():
    """The reference CPU encoder for the baseline arm: oracle/_ref if it was built, else the port."""
    from brotli_libs import REF_SO, Oracle, Ref
    if os.path.exists(REF_SO):
        return "reference", Ref()
    return "port", Oracle()


def cpu_time_one(lib, data):
    t = time.time()
    out = lib.compress(data, QUALITY, LGWIN)
    return time.time() - t, len(out)


def run_reference(args, rank, world):
    if rank != 0:
        return
    kind, lib = ref_lib()
    data = make_input(0)
    # one stream cannot use more than one core in the reference (single-threaded per state);
    # with N > 1 shards the host runs one encoder per shard on separate cores.
    n_streams = args.gpus
    cores = min(n_streams, os.cpu_count() or 1)
    datas = [data] if n_streams == 1 else [data] * n_streams   # same shape per shard; content reuse keeps setup short
    def step():
        if n_streams == 1:
            return cpu_time_one(lib, datas[0])[0]
        th, res = [], [0.0] * n_streams
        t0 = time.time()
        def work(i):
            lib.compress(datas[i], QUALITY, LGWIN)
        for i in range(n_streams):
            x = threading.Thread(target=work, args=(i,)); x.start(); th.append(x)
        for x in th:
            x.join()
        return time.time() - t0
    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = n_streams * WORKLOAD_BYTES * args.steps / total / 1e6
    line = {"impl": "reference", "metric": "encoder input MB/s (quality 5, lgwin 22, bit-exact)", "value": round(value, 2),
            "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * total / args.steps, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams": n_streams},
            "cpu_baseline": {"value": round(value, 2), "unit": "MB/s", "cores": cores, "kind": kind,
                             "sample": "the full %d-byte stream per step, %d stream(s) on %d core(s) (ctypes releases the GIL)" % (WORKLOAD_BYTES, n_streams, cores)},
            "e2e": {"value": round(value, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# dram__bytes_read.sum + dram__bytes_write.sum of k_walk's first launch on this workload (profiles/r01l_summary.md)
WALK_DRAM_GB = 30.61


_REAL_STDOUT = None


def _claim_stdout():
    """Everything any library prints on stdout (NCCL's version banner, ...) goes to stderr; the one JSON
    line is written to the real stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--bytes", type=int, default=WORKLOAD_BYTES, help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import brotli_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: brotli_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = brotli_b200.lib()
    data = make_input(rank, args.bytes)
    n = len(data)
    h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    d_in = h_in.cuda()
    cap = L.BrotliEncoderMaxCompressedSize(n) + 64
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()

    def gather_to_rank0(nbytes):
        """variable-size gather of the compressed shards (NCCL): sizes, then payloads."""
        if world == 1:
            return
        sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
        mine = torch.tensor([nbytes], dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(sizes, mine)
        if rank == 0:
            sz = sizes.tolist()
            bufs = [torch.empty(int(s), dtype=torch.uint8, device="cuda") for s in sz]
            reqs = [dist.irecv(bufs[r], src=r) for r in range(1, world)]
            for q in reqs:
                q.wait()
        else:
            dist.send(d_out[:nbytes], dst=0)

    def step_device():
        sz = C.c_size_t(cap)
        ok = L.BrotliB200CompressDevice(QUALITY, LGWIN, n, d_in.data_ptr(), C.byref(sz), d_out.data_ptr())
        assert ok, "BrotliB200CompressDevice failed"
        gather_to_rank0(sz.value)
        return sz.value

    def step_e2e():
        sz = C.c_size_t(cap)
        ok = L.BrotliEncoderCompress(QUALITY, LGWIN, 0, n, h_in.data_ptr(), C.byref(sz), h_out.data_ptr())
        assert ok, "BrotliEncoderCompress failed"
        return sz.value

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        outs = [fn() for _ in range(steps)]
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, outs

    for _ in range(args.warmup):
        out_size = step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dt, outs = timed(step_device, args.steps)
    st = brotli_b200.last_stats()
    sampler.stop_flag = True
    out_size = outs[-1]
    step_e2e()
    dt_e2e, outs2 = timed(step_e2e, args.steps)
    assert outs2[-1] == out_size

    # parity of this very run against the reference, outside the timed region (rank 0)
    parity = None
    cpu = None
    if rank == 0:
        kind, lib = ref_lib()
        t_cpu, ref_size = cpu_time_one(lib, data)
        want = lib.compress(data, QUALITY, LGWIN) if False else None
        got = bytes(h_out[:out_size].numpy().tobytes())
        t1 = time.time(); want = lib.compress(data, QUALITY, LGWIN); t_cpu2 = time.time() - t1
        parity = (got == want)
        t_best = min(t_cpu, t_cpu2)
        cpu = {"value": round(n / t_best / 1e6, 2), "unit": "MB/s", "cores": 1, "kind": kind,
               "sample": "the whole %d-byte stream, best of 2 runs, 1 thread (the reference encoder is single-threaded per stream)" % n}

    if rank == 0:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        value = world * n * args.steps / dt / 1e6
        e2e = world * n * args.steps / dt_e2e / 1e6
        # dominant kernel of the step: bytes it must move / its duration
        walk_ms, enc_ms = st["ms_walk"], st["ms_encode"]
        if walk_ms >= enc_ms:
            kname = "k_walk"
            algo = st["walk_bytes"] + 16.0 * st["total_cmds"] * st["walk_launches"] / max(1.0, st["lz77_iterations"])
            per_launch = algo / max(1.0, st["walk_launches"])
            dur = walk_ms / max(1.0, st["walk_launches"]) / 1e3
        else:
            kname = "k_encode_mb"
            per_launch = n + 16.0 * st["total_cmds"] + out_size
            dur = enc_ms / 1e3
        achieved = per_launch / dur / 1e9
        line = {"metric": "encoder input MB/s (quality 5, lgwin 22, bit-exact)", "value": round(value, 2), "unit": "MB/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1e3 * dt / args.steps, 2), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "input_bytes_per_gpu": n, "l2": "input (100 MB) and index (~3.5 GB) exceed the 126 MB L2",
                           "bit_exact_vs_reference": parity, "compressed_bytes": out_size},
                "clocks": sampler.summary(),
                "e2e": {"value": round(e2e, 2), "unit": "MB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": out_size},
                "gpu_launches": int(st["launches"]) * args.steps,
                "roofline": {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 6), "traffic": WALK_DRAM_GB if kname == "k_walk" else None,
                             "traffic_unit": "GB of DRAM read+write per launch (ncu --set full, first = dominant k_walk launch of this workload, profiles/r01l_summary.md)",
                             "algorithmic_gb_per_launch": round(per_launch / 1e9, 4),
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"},
                "cpu_baseline": cpu,
                "stages_ms": {k: round(st[k], 2) for k in ("ms_total", "ms_index", "ms_lz77", "ms_entropy", "ms_assemble", "ms_walk", "ms_encode")},
                "lz77": {"iterations": int(st["lz77_iterations"]), "block_runs": int(st["block_runs"]), "blocks": int(st["blocks"]),
                         "metablocks": int(st["metablocks"])}}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

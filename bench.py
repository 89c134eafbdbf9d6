#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 NTT engine (contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU NTT on host cores

Workload (BASELINE.json configs[1] / metric): forward Merge-NTT, Data64, N = 2^16, batch = 1024
polynomials PER GPU (weak scaling; 8 GPUs = configs[4], 8192 polynomials), single 60-bit prime
576460756061519873, X^N-1, in place, through the C ABI (gpuntt_b200_merge_ntt).  A "step" is one
call on the whole resident batch (512 MiB per GPU, > the 126 MB L2, so no L2 flush is needed).

One JSON line on stdout (rank 0).  `value` = polynomials transformed per second with data resident
in HBM (CUDA events, max over ranks); `e2e` = the same through gpuntt_b200_merge_ntt_host with
pinned HOST buffers, H2D + D2H inside the timed region; `roofline` = algorithmic bytes per launch of the
dominant kernel (fused2_kernel / fast_pass_kernel) / its live CUDA-event duration against the measured HBM copy
peak; `parity_gate` = the first step's output compared with the CPU reference before anything is timed (exit 3 on
a mismatch);
`cpu_baseline` = the reference's own NTTCPU<Data64>::ntt (oracle/_ref) on all host threads over a
bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOGN, BATCH, BITS = 16, 1024, 64
METRIC = "forward-NTT/s Data64 N=2^16 batch=1024"
UNIT = "NTT/s"
WORKLOAD = ("C2: Merge-NTT forward Data64 N=2^16 batch=1024 per GPU, single 60-bit prime "
            "576460756061519873, X^N-1, in place (GPU_NTT_Inplace)")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- CPU arms
def cpu_reference_rate(count, threads=None, seed=0):
    """(NTT/s, threads, kind) of the reference's CPU NTT over `count` polynomials of the workload."""
    from oracle import oracle as O
    R = O.ref()
    if R is not None:
        threads = threads or R.ref_hardware_threads()
        secs = R.ref_time_merge_ntt(LOGN, O.X_N_minus, BITS, count, threads, seed)
        return count / secs, threads, "reference"
    # oracle/_ref absent: time the C restatement instead (one thread per core, GIL released in ctypes)
    import concurrent.futures as cf
    import numpy as np
    threads = threads or os.cpu_count() or 1
    P = O.merge_params(LOGN, O.X_N_minus, BITS)
    x = O.example_input(P.modulus, count << LOGN, seed).reshape(count, -1)
    L = O.lib()

    def work(rows):
        for r in rows:
            L.ora_merge_ntt(r, LOGN, P.modulus, P.fwd, P.poly)
    parts = [list(x[i::threads]) for i in range(threads)]
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, parts))
    return count / (time.perf_counter() - t0), threads, "port"


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout under
    torchrun), so keep a private handle on the real stdout and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(text):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def reference_gpu_same_box():
    """The reference's OWN GPU kernels on this GPU, for the same workload: tools/bin/api_bench_reference is
    tools/api_bench.cu (a caller of the public GPU-NTT API) linked against the reference's GPU sources compiled for
    sm_100 (`make -C oracle refgpu`, built where /root/reference exists; the binary travels with the snapshot).
    Informational -- BASELINE.md section 4 item 4; the graded reference arm is the CPU one.  None when not built."""
    exe = os.path.join(ROOT, "tools", "bin", "api_bench_reference")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, "reference", "c2"], capture_output=True, text=True, timeout=120)
        for line in r.stdout.splitlines():
            if line.startswith("{"):
                d = json.loads(line)
                return {"value": d["ntt_per_s"], "unit": UNIT, "ms_per_step": d["ms"], "parity_vs_NTTCPU": d["parity_vs_NTTCPU"],
                        "what": "GPU_NTT_Inplace<Data64> of the reference's ntt.cu built for sm_100, same N/batch/prime, "
                                "CUDA events, data resident in HBM (tools/api_bench.cu)"}
    except Exception as e:  # noqa: BLE001 -- informational leg only
        return {"error": str(e)[:200]}
    return None


def workload_config(world):
    """`config` of the JSON line: the workload, identical for the GPU arm and the reference arm (what is specific to the GPU
    implementation -- the launch plan -- sits beside it under "plan")."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH,
            "l2": "inputs (512 MiB per GPU) larger than the 126 MB L2; no flush",
            "input": "seed-0 std::mt19937 + uniform_int_distribution stream of the reference examples, polynomials "
                     f"[{BATCH}*rank, {BATCH}*(rank+1)); timed steps re-run on the transformed data (values stay < p)",
            "partition": f"batch slices, {world} x {BATCH} polynomials, no collective"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    R = O.ref()
    threads = (R.ref_hardware_threads() if R is not None else os.cpu_count()) or 1
    per_step = max(64, 2 * threads)   # bounded sample of the 1024-polynomial batch per step
    for _ in range(args.warmup):
        cpu_reference_rate(per_step, threads)
    total_t, kind = 0.0, "port"
    for s in range(args.steps):
        rate, _, kind = cpu_reference_rate(per_step, threads, seed=s)
        total_t += per_step / rate
    value = per_step * args.steps / total_t
    sample = (f"{per_step} of the {BATCH} polynomials per step (the head of the examples' mt19937 stream, seeded with the step index), "
              f"NTTCPU<Data64>::ntt N=2^16 on {threads} host threads")
    emit_line(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.gpus),   # the same workload as the GPU arm, key for key
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample + " (reference CPU implementation on host cores; step = bounded sample)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.path = f"/tmp/bench_clocks_{os.getpid()}.csv"
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(dev)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples_under_load": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                util = float(p[4])
                if util < 50:
                    continue
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except ValueError:
                continue
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples_under_load=len(sm))
        return out


# ----------------------------------------------------------------------------- our arm
def seed0_slice(capi, p, rank):
    """Polynomials [BATCH * rank, BATCH * (rank + 1)) of the reference examples' seed-0 stream (SURVEY 8d:
    std::mt19937 gen(0) + uniform_int_distribution<uint64_t>(0, p-1), polynomial-major), as a numpy [BATCH, N] array.
    The distribution consumes a data-dependent number of draws, so rank r walks the stream from its start."""
    import numpy as np
    n = 1 << LOGN
    out = np.empty((rank + 1) * BATCH * n, dtype=np.uint64)
    capi.lib().gpuntt_b200_example_input(0, p, out.size, out.ctypes.data)
    return out[rank * BATCH * n:].reshape(BATCH, n)


def cpu_check_and_rate(x_host, y_gpu_host, count, threads=None):
    """bench.py's cpu_baseline leg: the reference's NTTCPU<Data64>::ntt over the first `count` polynomials of THIS run's
    input on all host threads -- timed (the baseline) and compared word for word with what the GPU produced for them in
    its first step (the parity gate).  Returns (NTT/s, threads, kind, mismatching words)."""
    import numpy as np
    from oracle import oracle as O
    R = O.ref()
    xin = np.ascontiguousarray(x_host[:count])
    want = np.empty_like(xin)
    if R is not None:
        threads = threads or R.ref_hardware_threads()
        secs = R.ref_time_merge_ntt_io(LOGN, O.X_N_minus, BITS, count, threads, xin.reshape(-1), want.reshape(-1))
        kind = "reference"
    else:
        import concurrent.futures as cf
        threads = threads or os.cpu_count() or 1
        P = O.merge_params(LOGN, O.X_N_minus, BITS)
        L = O.lib()
        want[:] = xin

        def work(rows):
            for r in rows:
                L.ora_merge_ntt(want[r], LOGN, P.modulus, P.fwd, P.poly)
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, [list(range(i, count, threads)) for i in range(threads)]))
        secs = time.perf_counter() - t0
        kind = "port"
    bad = int((want != y_gpu_host[:count]).sum())
    return count / secs, threads, kind, bad


def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gpu_ntt_b200 import capi
    from gpu_ntt_b200.params import NTTParameters, X_N_minus

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.lib()  # raises if the extension is missing

    P = NTTParameters(LOGN, X_N_minus, BITS)
    p = P.modulus
    h_tab = P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table)
    if args.corrupt_twiddle:                       # demonstration that the parity gate bites
        h_tab = h_tab.copy()
        h_tab[12345] ^= np.uint64(1)
    table = torch.from_numpy(h_tab.view(np.int64)).cuda()
    x_host = seed0_slice(capi, p, rank)            # SURVEY 8d input, this rank's batch slice
    data = torch.from_numpy(x_host.view(np.int64)).cuda()
    stream = torch.cuda.current_stream()

    def step():
        capi.ntt(data, table, p, LOGN, X_N_minus, stream=stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- first step + parity gate (before anything is timed)
    step()
    torch.cuda.synchronize()
    launches_per_step = lib.gpuntt_b200_last_launch_count()
    y_host = data.cpu().numpy().view(np.uint64)
    gate = {"checked_polynomials": 0, "mismatching_words": 0, "against": None}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        rate, threads, kind, bad = cpu_check_and_rate(x_host, y_host, min(args.cpu_sample, BATCH))
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"{min(args.cpu_sample, BATCH)} polynomials of this run's input (N=2^16, Data64, the examples' seed-0 "
                                  f"mt19937 stream), NTTCPU::ntt sharded over {threads} host threads"}
        gate = {"checked_polynomials": min(args.cpu_sample, BATCH), "mismatching_words": bad,
                "against": "oracle/_ref NTTCPU<Data64>::ntt" if kind == "reference" else "oracle/ntt_oracle.c"}
    else:
        # no CPU leg in this mode (multi-GPU ranks, --quick): cross-check the tuned kernels against the generic pass kernel
        # of the same library on the first 16 polynomials (two independent implementations; the CPU gate runs at N=1)
        chk = torch.from_numpy(np.ascontiguousarray(x_host[:16]).view(np.int64)).cuda()
        lib.gpuntt_b200_force_generic_path(1)
        capi.ntt(chk, table, p, LOGN, X_N_minus, stream=stream)
        lib.gpuntt_b200_force_generic_path(0)
        torch.cuda.synchronize()
        bad = int((chk.cpu().numpy().view(np.uint64) != y_host[:16]).sum())
        gate = {"checked_polynomials": 16, "mismatching_words": bad, "against": "generic pass kernel of the same library"}
    if gate["mismatching_words"]:
        sys.stderr.write(f"bench.py: PARITY GATE FAILED on rank {rank}: {gate}\n")
        sys.stderr.flush()
        os._exit(3)

    # ---- timed region: un-instrumented steps on the resident batch (values stay below p)
    warm = max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(warm):
        step()
    barrier()
    launches0 = lib.gpuntt_b200_total_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.gpuntt_b200_total_launch_count() - launches0
    # ---- second, instrumented loop: per-launch durations for the roofline (CUDA events around every launch)
    lib.gpuntt_b200_set_profiling(1)
    capi.profile_read()
    prof_steps = min(args.steps, 20)
    for _ in range(prof_steps):
        step()
    torch.cuda.synchronize()
    recs = capi.profile_read()
    lib.gpuntt_b200_set_profiling(0)
    # keep the same load running ~1 s more (untimed) so the 100 ms clock sampler sees it
    t_hold = time.perf_counter()
    while sampler is not None and not args.quick and time.perf_counter() - t_hold < 1.2:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * BATCH * args.steps / (ms * 1e-3)

    # live roofline of the dominant kernel
    pass_ms = [m for k, m in recs if k >= 1]
    npasses = max(1, len(pass_ms) // max(1, prof_steps))
    alg_bytes_per_launch = 2 * (1 << LOGN) * 8 * BATCH / npasses
    avg_pass_ms = sum(pass_ms) / max(1, len(pass_ms))
    peak, peak_src = measured_peak()
    achieved = alg_bytes_per_launch / (avg_pass_ms * 1e-3) / 1e9 if pass_ms else None
    kernel = ("fused2_kernel (one launch per step: strided stages 0-7 and contiguous stages 8-15 chained through the L2)"
              if npasses == 1 else
              "fast_pass_kernel (%d launches per step: strided stages 0-7, contiguous stages 8-15)" % npasses)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        ent = tj.get("fused2_kernel" if npasses == 1 else "fast_pass_kernel")
        if ent:
            traffic = ent["dram_bytes_per_launch"]        # from one ncu --set full capture (not measured in this run)
            traffic_src = {k: ent.get(k) for k in ("capture", "commit", "kernel")}
    except Exception:
        pass
    # informational second roof: the 32-bit integer multiplier pipe.  One 64-bit Shoup butterfly costs
    # 28 issue cycles of that pipe per warp and SM sub-partition (4 IMAD.WIDE + 1 IMAD.HI at 4, 4 IMAD at 2 --
    # profiles/r1_pipe_probes2.txt); an N = 2^16 cyclic transform has 16 * 2^15 butterflies of which the
    # twiddle-1 ones of the first three stages (2^15 + 2^14 + 2^13) need no multiply.
    muls_per_ntt = LOGN * (1 << (LOGN - 1)) - ((1 << 15) + (1 << 14) + (1 << 13))
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    clk = (clocks or {}).get("sm_mhz") or 1965.0
    int_roof = sms * 4 * clk * 1e6 / (28.0 * muls_per_ntt / 32.0)

    # end to end: pinned host buffers -> H2D -> NTT -> D2H through the host-buffer C-ABI entry point
    import ctypes as C
    h_in = torch.empty((BATCH, 1 << LOGN), dtype=torch.int64).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    h_in.copy_(torch.from_numpy(x_host.view(np.int64)))
    desc = capi.MergeDesc(BITS, 0, capi.FORWARD, LOGN, capi.PerPolynomial, X_N_minus, BATCH, 0,
                          h_in.data_ptr(), h_out.data_ptr(), None, p, 0, None, None, stream.cuda_stream)

    def e2e_step():
        capi.check(lib.gpuntt_b200_merge_ntt_host(C.byref(desc), h_tab.ctypes.data, h_tab.size))
    e2e_step()
    if int((h_out.numpy().view(np.uint64) != y_host).sum()):
        sys.stderr.write("bench.py: PARITY GATE FAILED: the host-buffer entry point disagrees with the device-resident call\n")
        os._exit(3)
    e2e_steps = 1 if args.quick else max(1, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()      # synchronises its stream before returning
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    nbytes = BATCH * (1 << LOGN) * 8

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": workload_config(world),
        "plan": {"passes": capi.describe_plan(LOGN, BITS).strip(), "launches_per_step": launches_per_step},
        "parity_gate": gate,
        "clocks": clocks,
        "e2e": {"value": world * BATCH * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nbytes + h_tab.nbytes,
                "d2h_bytes_per_step": nbytes, "steps": e2e_steps,
                "api": "gpuntt_b200_merge_ntt_host (pinned host in/out, copies + kernels + sync per step)"},
        "gpu_launches": int(launches),
        "int_mul_roof": {"ntt_per_s_per_gpu": int_roof, "frac": (value / world) / int_roof,
                         "model": "28 multiplier-pipe issue cycles per warp-butterfly, %d multiplying butterflies per NTT, "
                                  "%d SMs x 4 sub-partitions at %.0f MHz" % (muls_per_ntt, sms, clk)},
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src, "launches_per_step": npasses,
                     "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_pass_ms,
                     "timing": "CUDA events around every launch in a second, instrumented loop (not the timed region)",
                     "frac_of_8TBps_nominal": (achieved / 8000.0) if achieved else None},
    }
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
        out["reference_gpu_same_box"] = reference_gpu_same_box()
    if rank == 0:
        emit_line(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="kernel-timed region only (for runs under ncu)")
    ap.add_argument("--corrupt-twiddle", action="store_true", help="flip one bit of the root table: the parity gate must exit non-zero")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()

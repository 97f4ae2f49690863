#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 shooting loop (BASELINE.json: visibility rays/sec and
seconds-to-converge on church.tri), contract as described in the build brief.

A "step" is one complete run of the hot path on the workload: rewind the radiosity state (device
copy), flush L2, then shoot to convergence.  `value` is reference-equivalent visibility rays per
second (one ray = one RayHitted call of the reference) with all inputs
resident in HBM; `e2e` is the same metric through the C ABI with HOST buffers (state upload from the
host TriStore image + loop + download inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, Model kwargs, Context kwargs, description)
    "church": ("church", {}, {}, "model/church.tri -triangle 1000000, default options, to convergence"),
    "room": ("room", {}, {}, "model/room.tri -converge 1200 -delta_ff 0.000005, first 200 shots"),
    "box": ("box", {}, {}, "model/box.tri default options, first 500 shots"),
    "blocks": ("blocks", {}, {}, "model/blocks.tri frozen mesh, to convergence"),
    # seeded synthetic rooms (hw-radiosity_b200/synth.py), refined with the exhaustive PrePartition variant
    # -converge 1 keeps candidates alive for the fixed shot count (the small ceiling light alone would end after 2 shots)
    "synth1m": ("synth:700000", {"prepartition": 2, "tri_limit": 4000000, "max_tri": 4000000}, {"converge": 1.0},
                "synthetic room, seed 1, ~1.4 M receivers, -converge 1, first 200 shots"),
    "synth10m": ("synth:8100000", {"prepartition": 2, "tri_limit": 26000000, "max_tri": 26000000}, {"converge": 1.0},
                 "synthetic room, seed 1, 11.6 M receivers, -converge 1, fixed 1000 shots"),
}
SHOT_CAP = {"room": 200, "box": 500, "synth1m": 200, "synth10m": 1000}
ALGO_BYTES_PER_RECEIVER = 160


def unpack_model(name, tmp):
    path = os.path.join(tmp, name.replace(":", "_") + ".tri")
    with open(path, "wb") as f:
        if name.startswith("synth:"):
            from hw_radiosity_b200 import synth
            f.write(synth.room_scene(int(name.split(":")[1]), seed=1))
        else:
            f.write(gzip.open(os.path.join(ROOT, "tests", "golden", "models", name + ".tri.gz")).read())
    return path


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, idx in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > idx and r[idx].lower() == "active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_run(tri_path, shots, threads, build="ref_harness"):
    """The reference's own CPU implementation (oracle/_ref/ref_harness = unmodified sources, -O2
    -ffp-contract=off -fopenmp; ref_harness_shipped = the reference Makefile's own flags, /root/reference/Makefile:7)
    timed around its shot loop; falls back to the oracle port."""
    from oracle import pyoracle as po
    harness = po.ref_binary(build)
    if harness:
        args = [harness, "-jobs", str(threads), "-o", os.path.join(os.path.dirname(tri_path), "ref_out"), tri_path, "--json"]
        if shots is not None:
            args += ["--shots", str(shots)]
        out = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode(errors="replace")
        info = json.loads([l for l in out.splitlines() if l.startswith("HARNESS_JSON")][0][len("HARNESS_JSON "):])
        return {"kind": "reference", "rays": info["rays"], "seconds": info["loop_s"], "shots": info["shots"]}
    import hw_radiosity_b200 as hr
    m = hr.Model(tri_path)
    s = m.scene
    sc = po.Scene(m.triangles.copy(), m.n0, m.nodes.copy(), m.tri_list.copy(), s.G0, s.G1, s.min_grid_len, s.grid_num,
                  threads=threads)
    t = time.perf_counter()
    seq = sc.run(shots)
    dt = time.perf_counter() - t
    return {"kind": "port", "rays": sc.stats["rays"], "seconds": dt, "shots": len(seq)}


def cpu_port_run(model, shots, threads, converge=1200.0):
    """Oracle port (oracle/radiosity_oracle.c, OpenMP) on a scene the reference's static limits cannot hold."""
    from oracle import pyoracle as po
    s = model.scene
    sc = po.Scene(model.triangles.copy(), model.n0, model.nodes.copy(), model.tri_list.copy(), s.G0, s.G1, s.min_grid_len,
                  s.grid_num, threads=threads, converge=converge)
    t = time.perf_counter()
    seq = sc.run(shots)
    dt = time.perf_counter() - t
    return {"kind": "port", "rays": sc.stats["rays"], "seconds": dt, "shots": len(seq)}


def run_reference_arm(args):
    """The reference's own CPU implementation on this box's host cores, all threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    model, mkw, ckw, desc = WORKLOADS[args.workload]
    with tempfile.TemporaryDirectory() as tmp:
        tri = unpack_model(model, tmp)
        if model.startswith("synth:"):
            import hw_radiosity_b200 as hr                      # host front end only (no GPU needed)
            mdl = hr.Model(tri, **mkw)
            run = lambda shots: cpu_port_run(mdl, shots if shots else SHOT_CAP[args.workload], threads, ckw.get("converge", 1200.0))
        else:
            run = lambda shots: cpu_reference_run(tri, shots, threads)
        # bounded sample: size a step so that (steps + warmup) steps stay within ~2.5 minutes
        probe = run(3 if model.startswith("synth:") else 10)
        per_shot = max(probe["seconds"] / max(probe["shots"], 1), 1e-6)
        total = SHOT_CAP.get(args.workload)
        budget = 150.0 / (args.steps + args.warmup)
        shots = int(max(2, min(budget / per_shot, total or 10 ** 9)))
        if total is None and budget / per_shot > 400:
            shots = None                                        # the whole run fits: go to convergence
        for _ in range(args.warmup):
            run(shots)
        rays, secs, nshots, kind = 0, 0.0, 0, "reference"
        for _ in range(args.steps):
            r = run(shots)
            rays += r["rays"]; secs += r["seconds"]; nshots = r["shots"]; kind = r["kind"]
    value = rays / secs
    sample = f"{model} first {nshots} shots per step" if shots else f"{model} to convergence ({nshots} shots) per step"
    line = {
        "impl": "reference", "metric": "visibility_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "repo model (tests/golden/models, gzip of the reference's model/*.tri)" if not model.startswith("synth:") else "synthetic (seeded generator)",
        "config": {"workload": desc, "sample": sample, "threads": threads},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))



def golden_entry(name):
    """Committed golden of a repo model run to convergence by the UNMODIFIED reference (tests/golden/*.json)."""
    gdir = os.path.join(ROOT, "tests", "golden")
    try:
        g = json.load(open(os.path.join(gdir, "golden.json")))["models"].get(name, {}).get("full")
        if g:
            g = dict(g, shooters=[s[0] for s in g["shooters"]])
        else:
            g = json.load(open(os.path.join(gdir, "golden_full.json")))["models"].get(name)
        return g
    except (OSError, ValueError, KeyError):
        return None


def parity_block(ctx, model, rank, world, gold, done):
    """After an untimed run to convergence: gather accB / deltaB over the ranks and compare with the golden of the
    unmodified reference - SHA-256 over the live slots, and the shooter sequence.  None without a golden."""
    import hashlib
    import numpy as np
    from hw_radiosity_b200 import dist as rdist
    state = rdist.gather_state(ctx, rank, world)
    if gold is None:
        return None
    if rank != 0:
        return {}
    accB, deltaB = state
    live = model.triangles["parent"] >= 0        # the frozen mesh's receivers, in slot order (the order the goldens hash)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return {"accB_sha_ok": sha(accB[live]) == gold["accB_sha"], "deltaB_sha_ok": sha(deltaB[live]) == gold["deltaB_sha"],
            "shooters_ok": done == gold["shots"] and ctx.shooter_log() == list(gold["shooters"]),
            "against": "tests/golden (unmodified reference, -ffp-contract=off)"}


def extra_workload(name, rank, world, local_rank, barrier, max_over_ranks, flush):
    """Device time of another BASELINE config at the current N (2 timed runs after 1 warm-up, L2 flushed, state rewound
    on the device), with the parity flags of an untimed run where a golden exists."""
    import hw_radiosity_b200 as hr
    from hw_radiosity_b200 import dist as rdist
    model_name, mkw, ckw, desc = WORKLOADS[name]
    cap = -1 if name in ("box", "room") else SHOT_CAP.get(name, -1)
    t_setup = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        model = hr.Model(unpack_model(model_name, tmp), **mkw)
    ctx = rdist.sharded_context(model, rank, world, local_rank, **ckw)
    setup_s = time.perf_counter() - t_setup
    ms = []
    done = 0
    for i in range(3):
        ctx.rewind()
        flush.fill_(i + 1)
        import torch
        torch.cuda.synchronize()
        barrier()
        before = ctx.device_ms
        done, conv = ctx.step(cap)
        ms.append(max_over_ranks(ctx.device_ms - before))
    gold = golden_entry(model_name) if cap < 0 else None
    par = parity_block(ctx, model, rank, world, gold, done)
    sizes = ctx.sizes()
    out = {"workload": desc if cap >= 0 else desc.split(",")[0].replace(" default options", "") + ", default options, to convergence", "device_ms": min(ms[1:]),
           "shots": done, "us_per_shot": min(ms[1:]) * 1e3 / max(done, 1), "receivers_per_rank": sizes["receivers"],
           "setup_s": round(setup_s, 2), "parity": par}
    if gold:
        out["reference_equivalent_rays"] = gold["rays"]
        out["rays_per_s"] = gold["rays"] / (min(ms[1:]) * 1e-3)
    # bytes side of the SURVEY 8(d) bound with EVERY receiver of the frozen mesh counted as shaded in every shot (saturated
    # receivers are not subtracted: an upper bound on the algorithmic bytes, so the fraction is an upper bound too)
    leaves = int((model.triangles["parent"] >= 0).sum())
    hbm_gbs = measured_peaks()[0] * world
    out["roofline_bytes_side"] = {"receivers": leaves, "bound_us_per_shot": ALGO_BYTES_PER_RECEIVER * leaves / (hbm_gbs * 1e9) * 1e6,
                                  "frac_upper_bound": ALGO_BYTES_PER_RECEIVER * leaves * max(done, 1) / (hbm_gbs * 1e9) / (min(ms[1:]) * 1e-3)}
    ctx.close()
    model.close()
    barrier()
    return out


def split_workload():
    """BASELINE config 3: blocks.tri with adaptive splitting enabled, -adapt_area 5, to convergence (one GPU)."""
    import hw_radiosity_b200 as hr
    with tempfile.TemporaryDirectory() as tmp:
        m = hr.Model(unpack_model("blocks", tmp), prepartition=False, tri_limit=1000000, sample_area=-1.0)
    best = None
    for rep in range(3):
        ctx = hr.Context(m, split_enabled=1, area_limit=5.0, sample_area=m.sample_area, tri_limit=1000000)
        done, conv = ctx.step(-1)
        st = ctx.stats()
        rec = {"workload": "model/blocks.tri -adapt_area 5, adaptive splitting enabled, to convergence", "device_ms": ctx.device_ms,
               "shots": done, "converged": conv, "slots": ctx.trinum, "splits": st["splits"], "rays": st["rays"],
               "kernel_launches": st["kernel_launches"],
               "parity": "statistical (own sampling RNG): tests/test_split.py; exact goldens without sampling"}
        ctx.close()
        if rep and (best is None or rec["device_ms"] < best["device_ms"]):
            best = rec
    return best


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import hw_radiosity_b200 as hr
    from hw_radiosity_b200 import dist as rdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    # stdout carries exactly one JSON line: library chatter (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local_rank = rdist.init("nccl")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
    model_name, mkw, ckw, desc = WORKLOADS[args.workload]
    cap = SHOT_CAP.get(args.workload, -1)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with tempfile.TemporaryDirectory() as tmp:
        tri = unpack_model(model_name, tmp)
        model = hr.Model(tri, **mkw)
        # reference-equivalent ray count of one step (deterministic): one statistics run on rank 0, untimed
        info = [None]
        if rank == 0:
            sctx = hr.Context(model, collect_stats=1, device=local_rank, **ckw)
            sctx.step(cap)
            st = sctx.stats()
            info[0] = st
            sctx.close()
        if world > 1:
            dist.broadcast_object_list(info, src=0)
        st = info[0]
        rays_per_step, shots_per_step, recv_shaded = st["rays"], st["shots"], st["receivers"]
        walk = {k: st[k] for k in ("traced_rays", "leaf_visits", "descent_steps", "tri_tests", "tri_accepts", "visible", "calff")}

        ctx = rdist.sharded_context(model, rank, world, local_rank, **ckw)
        sizes = ctx.sizes()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

        def one_step():
            ctx.rewind()
            flush.fill_(rank + 1)
            torch.cuda.synchronize()
            barrier()
            before = ctx.device_ms
            done, conv = ctx.step(cap)
            assert done == shots_per_step, (done, shots_per_step)
            return max_over_ranks(ctx.device_ms - before)       # device time, max over ranks

        # clocks / throttle reasons are sampled from the first warm-up step to the end of the timed region (the timed
        # region alone lasts tens of milliseconds, shorter than one nvidia-smi sampling period)
        with ClockSampler(local_rank) as clk:
            for _ in range(max(args.warmup, 3)):
                one_step()
            l0 = ctx.stats()["kernel_launches"]
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            dev_ms = [one_step() for _ in range(args.steps)]
            torch.cuda.synchronize()
            barrier()
            wall = time.perf_counter() - t0
            launches = ctx.stats()["kernel_launches"] - l0
            if wall < 0.5:                      # keep the GPU under the same load until the sampler has a few readings
                t_end = time.perf_counter() + 0.6
                while time.perf_counter() < t_end:
                    one_step()
        ms_per_step = sum(dev_ms) / len(dev_ms)
        value = rays_per_step / (ms_per_step / 1e3)

        # per-kernel-class durations, live, CUDA events on the launching stream (collective: all ranks launch alike)
        ctx.rewind()
        flush.fill_(7)
        torch.cuda.synchronize()
        barrier()
        kt = ctx.profile(cap)
        ctx.rewind()

        # end-to-end through the C ABI with host buffers: state upload + loop + download
        host_out = ctx.download()            # caller-owned host buffers, reused every step

        phases = [0.0, 0.0, 0.0]

        def e2e_step():
            barrier()
            t = time.perf_counter()
            ctx.reset(model)                 # H2D: accB/deltaB of this rank's receivers from the host TriStore image
            t1 = time.perf_counter()
            ctx.step(cap)
            t2 = time.perf_counter()
            ctx.download(out=host_out)       # D2H: accB/deltaB of this rank's receivers into the host buffers
            t3 = time.perf_counter()
            for k, d in enumerate((t1 - t, t2 - t1, t3 - t2)):
                phases[k] += d
            return max_over_ranks(t3 - t)
        for _ in range(2):
            e2e_step()
        phases[:] = [0.0, 0.0, 0.0]
        e2e_s = sum(e2e_step() for _ in range(args.steps)) / args.steps
        e2e_phases = {k: round(v / args.steps * 1e3, 3) for k, v in zip(("upload_ms", "loop_ms", "download_ms"), phases)}
        state_bytes = sizes["receivers"] * 48

        # parity of THIS run's configuration (all N ranks, real peer mailboxes): one untimed run to convergence, state
        # gathered over the ranks, compared with the golden of the unmodified reference
        ctx.rewind()
        c0 = ctx.stats()
        pdone, _ = ctx.step(cap)
        parity = parity_block(ctx, model, rank, world, golden_entry(model_name) if cap < 0 else None, pdone)
        c1 = ctx.stats()
        counters = {k: c1[k] - (c0[k] if k == "shots" else 0) for k in ("shots", "ff_rows_filled", "rounds", "shots_scheduled", "rounds_reapplied")}

        # CPU baseline on this box's host cores (rank 0, N=1 only): the unmodified reference, bounded sample
        cpu = cpu_shipped = None
        threads = os.cpu_count() or 1
        if rank == 0 and world == 1:
            if model_name.startswith("synth:"):
                cpu = cpu_port_run(model, min(cap, 3), threads, ckw.get("converge", 1200.0))   # the reference's MaxTri cannot hold this scene
            else:
                cpu_shots = None if args.workload in ("church", "blocks") else min(cap, 40)
                cpu = cpu_reference_run(tri, cpu_shots, threads)
                try:        # the same loop built with the reference Makefile's own flags (FMA contraction allowed)
                    cpu_shipped = cpu_reference_run(tri, cpu_shots, threads, build="ref_harness_shipped")
                except Exception:
                    cpu_shipped = None
        barrier()
        ctx.close()

        # where a round's time goes (untimed, diagnostics build of the same kernels: %globaltimer stamps at the start / end
        # of every kernel of a round, RADK_DEBUG_PREDICT): shows the chain running beside the form factors, the exchanges
        timeline = None
        try:
            os.environ["RADK_DEBUG_PREDICT"] = "1"
            dctx = rdist.sharded_context(model, rank, world, local_rank, **ckw)
            os.environ.pop("RADK_DEBUG_PREDICT", None)
            for _ in range(3):
                dctx.rewind()
                barrier()
                dctx.step(cap)
            timeline = dctx.round_timeline()
            barrier()
            dctx.close()
        except Exception as e:          # diagnostics never take the headline line down
            timeline = {"error": f"{type(e).__name__}: {e}"[:200]}
        finally:
            os.environ.pop("RADK_DEBUG_PREDICT", None)
        model.close()

        # the other BASELINE configs at this N (device time + parity flags), after the headline measurement
        extras = {}
        if args.workload == "church" and not args.no_extra:
            for wl in args.extra.split(","):
                if not wl:
                    continue
                t_x = time.perf_counter()
                try:
                    if wl == "blocks_split":
                        extras[wl] = split_workload() if (rank == 0 and world == 1) else {"skipped": "adaptive splitting runs on one GPU"}
                    else:
                        extras[wl] = extra_workload(wl, rank, world, local_rank, barrier, max_over_ranks, flush)
                except Exception as e:      # an extra never takes the headline line down
                    extras[wl] = {"error": f"{type(e).__name__}: {e}"[:300]}
                extras[wl]["wall_s_total"] = round(time.perf_counter() - t_x, 1)

    hbm_gbs, peak_src, sm_max = measured_peaks()
    hbm_gbs *= world               # aggregate peaks of the N GPUs the job runs on
    kprof = {}                     # per-kernel ncu figures of the committed captures (1 GPU): issue slots, FP32 pipe, active lanes, DRAM bytes
    try:
        kprof = json.load(open(os.path.join(ROOT, "profiles", "r02_kernels.json"))).get(args.workload, {})
    except (OSError, ValueError):
        pass
    # T_shot >= max(160 B x shaded receivers / HBM peak, counted FP32 ops / FP32 peak); roofline.frac is
    # that bound over the MEASURED time of the whole step (every kernel of the loop, not the dominant one alone).
    ff_ms, ff_n = kt["ms"]["form_factor"], max(kt["launches"]["form_factor"], 1)
    recv_per_shot = recv_shaded / max(shots_per_step, 1)
    algo_bytes_per_step = ALGO_BYTES_PER_RECEIVER * recv_shaded
    flops = (35 * walk["calff"] + 12 * walk["leaf_visits"] + 6 * walk["descent_steps"] + 22 * walk["leaf_visits"]
             + 14 * walk["tri_tests"] + 24 * walk["tri_accepts"])
    p_fp32 = 148 * 128 * 2 * sm_max * 1e6 * world
    t_bytes, t_flops = algo_bytes_per_step / (hbm_gbs * 1e9), flops / p_fp32          # seconds per step at the two roofs
    step_s = ms_per_step * 1e-3
    if t_bytes >= t_flops:
        roof = {"bound": "hbm", "achieved": algo_bytes_per_step / step_s / 1e9, "peak": hbm_gbs, "unit": "GB/s"}
    else:
        roof = {"bound": "fp32", "achieved": flops / step_s / 1e12, "peak": p_fp32 / 1e12, "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    dom = max(kt["ms"], key=lambda k: kt["ms"][k])
    dom_name = {"form_factor": "ff_batch_kernel (K3a: culls + form factor + octree walk)", "update": "chain/apply/commit kernels (K3b)",
                "select": "predict_kernel (K1+K2)"}[dom]
    traffic = kprof.get(dom, {}).get("dram_bytes_per_launch") if world == 1 else None
    roof.update({
        "traffic": traffic, "kernel": dom_name, "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
        "definition": "SURVEY 8(d) whole-step fraction: max(160 B x shaded receivers / HBM peak, counted FP32 ops / FP32 peak) / measured device time of the step",
        "bound_us_per_shot": max(t_bytes, t_flops) / max(shots_per_step, 1) * 1e6,
        "measured_us_per_shot": ms_per_step * 1e3 / max(shots_per_step, 1),
        "hbm_side": {"algorithmic_bytes_per_step": algo_bytes_per_step, "seconds_at_peak": t_bytes},
        "fp32_side": {"flops_per_step_counted": flops, "seconds_at_peak": t_flops, "peak_tflops": p_fp32 / 1e12,
                      "achieved_tflops": flops / step_s / 1e12},
        "algorithmic_bytes_per_launch": algo_bytes_per_step / ff_n,
        "kernel_ms_per_step": kt["ms"], "kernel_launches_per_step": kt["launches"], "shots_profiled": kt["shots"],
        "kernels_ncu": kprof or None,
        "note": "church is neither HBM nor FP32-pipe bound in practice: the walk is issue / latency bound (kernels_ncu: issue slots, "
                "active lanes), the receiver update is a chain of small kernels; the fraction above is the north-star's definition"})
    runs = max(counters["rounds"], 1)
    line = {
        "metric": "visibility_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32",
        "data": "repo model (tests/golden/models, gzip of the reference's model/*.tri)" if not model_name.startswith("synth:") else "synthetic (seeded generator)",
        "config": {"workload": desc, "receivers_per_rank": sizes["receivers"], "unique_vertices_per_rank": sizes["unique_vertices"],
                   "parallelism": f"receiver shards x{world}" if world > 1 else "single GPU",
                   "shots_per_step": shots_per_step, "reference_equivalent_rays_per_step": rays_per_step,
                   "traced_rays_per_step": walk["traced_rays"], "l2": "flushed between steps (256 MiB write), state rewound on device",
                   "timing": "per-step CUDA-event windows around radk_ctx_step on its stream, mean of K steps",
                   "parity": "bit-exact mode (-fmad=false)"},
        "round_timeline_us": timeline,
        "seconds_to_converge": ms_per_step / 1e3 if cap < 0 else None,
        "wall_s_timed_region": wall,
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "e2e": {"value": rays_per_step / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": state_bytes,
                "d2h_bytes_per_step": state_bytes, "seconds_per_step": e2e_s, **e2e_phases},
        "roofline": roof,
        "parity": parity,
        "speculation": {"ff_rows_filled_per_shot": counters["ff_rows_filled"] / max(counters["shots"], 1),
                        "shots_scheduled_per_shot_committed": counters["shots_scheduled"] / max(counters["shots"], 1),
                        "rounds_reapplied_frac": counters["rounds_reapplied"] / runs,
                        "note": "rows filled = form-factor rows computed; every committed shot needs one, the excess was evicted unused"},
        "workloads": extras or None,
    }
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["rays"] / cpu["seconds"], "unit": "rays/s", "cores": threads, "kind": cpu["kind"],
                                "sample": f"{model_name}, {cpu['shots']} shots, shot loop only ({cpu['seconds']:.2f} s)",
                                "seconds": cpu["seconds"], "build": "-O2 -ffp-contract=off (the parity oracle)"}
        if cpu_shipped is not None:
            line["cpu_baseline"]["shipped_flags"] = {
                "value": cpu_shipped["rays"] / cpu_shipped["seconds"], "seconds": cpu_shipped["seconds"], "shots": cpu_shipped["shots"],
                "build": "-O2 -march=native -msse -msse2 (reference Makefile:7; FMA contraction allowed, results differ in the last bits)"}
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="church", choices=sorted(WORKLOADS))
    ap.add_argument("--extra", default="box,room,blocks_split,synth10m",
                    help="other BASELINE configs reported in the `workloads` block of the church line (device ms + parity)")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

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
def cpu_reference_run(tri_path, shots, threads):
    """The reference's own CPU implementation (oracle/_ref/ref_harness = unmodified sources, -O2
    -ffp-contract=off -fopenmp) timed around its shot loop; falls back to the oracle port."""
    from oracle import pyoracle as po
    harness = po.ref_binary("ref_harness")
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

        # CPU baseline on this box's host cores (rank 0, N=1 only): the unmodified reference, bounded sample
        cpu = None
        threads = os.cpu_count() or 1
        if rank == 0 and world == 1:
            if model_name.startswith("synth:"):
                cpu = cpu_port_run(model, min(cap, 3), threads, ckw.get("converge", 1200.0))   # the reference's MaxTri cannot hold this scene
            else:
                cpu_shots = None if args.workload in ("church", "blocks") else min(cap, 40)
                cpu = cpu_reference_run(tri, cpu_shots, threads)
        barrier()

    hbm_gbs, peak_src, sm_max = measured_peaks()
    hbm_gbs *= world               # aggregate peaks of the N GPUs the job runs on
    traffic_rec = {}               # DRAM bytes per launch of the dominant kernel, from the committed ncu capture (1 GPU)
    try:
        traffic_rec = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(args.workload, {})
    except (OSError, ValueError):
        pass
    # dominant kernel: the visibility / form-factor kernel (K3a: ff_batch_kernel).  Algorithmic bytes per launch =
    # 160 B x shaded receivers of the shots whose form factors the launch produced.
    ff_ms, ff_n = kt["ms"]["form_factor"], max(kt["launches"]["form_factor"], 1)
    upd_ms = kt["ms"]["update"]
    recv_per_shot = recv_shaded / max(shots_per_step, 1)
    algo_bytes_per_step = ALGO_BYTES_PER_RECEIVER * recv_shaded
    dom = "ff_batch_kernel (K3a: culls + form factor + octree walk)" if ff_ms >= upd_ms else "shots_kernel (K3b+K1)"
    dom_ms = max(ff_ms, upd_ms)
    traffic = traffic_rec.get("dram_bytes_per_launch") if world == 1 and dom.startswith(traffic_rec.get("kernel", "?")) else None
    achieved = algo_bytes_per_step / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # FP32 side of the SURVEY section 8(d) roofline, counted per traced ray by the device itself
    flops = (35 * walk["calff"] + 12 * walk["leaf_visits"] + 6 * walk["descent_steps"] + 22 * (walk["leaf_visits"] - walk["tri_accepts"] * 0)
             + 14 * walk["tri_tests"] + 24 * walk["tri_accepts"])
    p_fp32 = 148 * 128 * 2 * sm_max * 1e6 * world
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
        "seconds_to_converge": ms_per_step / 1e3 if cap < 0 else None,
        "wall_s_timed_region": wall,
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "e2e": {"value": rays_per_step / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": state_bytes,
                "d2h_bytes_per_step": state_bytes, "seconds_per_step": e2e_s, **e2e_phases},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                     "traffic": traffic, "kernel": dom, "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
                     "algorithmic_bytes_per_launch": algo_bytes_per_step / ff_n,
                     "note": "achieved = 160 B x shaded receivers of the step / summed launch time of the dominant kernel "
                             "(CUDA events per launch, profiled run); church is FP32-issue/latency bound, not HBM bound "
                             "(SURVEY 8d): see fp32 and shot_roofline",
                     "kernel_ms_per_step": kt["ms"], "kernel_launches_per_step": kt["launches"], "shots_profiled": kt["shots"]},
        "shot_roofline": {"definition": "SURVEY 8(d): T_shot >= max(160 B x active receivers / HBM peak, counted FP32 ops / FP32 peak)",
                          "bound_us_per_shot": max(ALGO_BYTES_PER_RECEIVER * recv_per_shot / (hbm_gbs * 1e9), flops / max(shots_per_step, 1) / p_fp32) * 1e6,
                          "measured_us_per_shot": ms_per_step * 1e3 / max(shots_per_step, 1),
                          "frac": max(ALGO_BYTES_PER_RECEIVER * recv_per_shot / (hbm_gbs * 1e9), flops / max(shots_per_step, 1) / p_fp32) / (ms_per_step * 1e-3 / max(shots_per_step, 1))},
        "fp32": {"flops_per_step_counted": flops, "achieved_tflops": flops / (ms_per_step * 1e-3) / 1e12,
                 "peak_tflops": p_fp32 / 1e12, "frac": flops / (ms_per_step * 1e-3) / p_fp32},
    }
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["rays"] / cpu["seconds"], "unit": "rays/s", "cores": threads, "kind": cpu["kind"],
                                "sample": f"{model_name}, {cpu['shots']} shots, shot loop only ({cpu['seconds']:.2f} s)",
                                "seconds": cpu["seconds"]}
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    ctx.close()
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

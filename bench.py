#!/usr/bin/env python
"""bench.py -- tracer node-level updates/s of the batched do_oce_adv_tra step on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      CPU restatement on the host cores

One "step" = one call of do_oce_adv_tra for the whole tracer batch (T+S, MFCT + QR4C + FCT),
including its two internal halo exchanges.  Default workload = BASELINE.json
config 4 scaled to the GPU count: the 1733x1733x70 (3.0M-node) synthetic mesh at N=8, and the same
node density per GPU below that (613x613 at N=1), METIS-partitioned -> weak scaling.

`value`   : device-resident inputs, CUDA events on the library's compute stream, max over ranks.
`e2e`     : the same step through the C ABI with HOST buffers (pinned), H2D/D2H inside the timing.
`roofline`: SURVEY 8d algorithmic bytes of the step / summed kernel time, against MEASURED_PEAKS.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "tracer node-level updates/sec (nod2D x nlev x ntracer / s)"
UNIT = "updates/s"
PEAK_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


# ------------------------------------------------------------------------------------------------
def workload_mesh(name: str, ngpus: int):
    """Global mesh of the named workload."""
    from fesom2_b200 import mesh as M
    if name == "cfg4":      # BASELINE config 4, per-GPU share of the 3.0M x 70 mesh
        side = int(round(1733 * math.sqrt(ngpus / 8.0)))
        return M.synth_mesh(side, side, nl=71), f"synthetic {side}x{side} lon-lat mesh, 70 layers (config 4: 3.0M nodes at 8 GPUs, same nodes/GPU below)"
    if name == "cfg4-full":  # strong-scaling variant: the whole 3.0M mesh at any N
        return M.synth_mesh(1733, 1733, nl=71), "synthetic 1733x1733 (3.0M-node) mesh, 70 layers (config 4, strong scaling)"
    if name in ("core2", "recom30"):   # configs 3 / 5
        return M.synth_mesh(357, 356, nl=48), "synthetic CORE2-sized 357x356 mesh (127k nodes), 47 layers"
    if name == "pi":
        return M.load_npz_mesh(os.path.join(ROOT, "tests", "golden", "mesh_pi.npz")), "pi test mesh (3140 nodes, 47 layers)"
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes(L, N, E, T, B):
    """8 L [ B (12 N + 4 E) + (3 T + 7 N) ] per rank and step."""
    return 8.0 * L * (B * (12.0 * N + 4.0 * E) + (3.0 * T + 7.0 * N))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.p = None
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_arm(args, g, desc, steps, warmup, ntr):
    """The CPU restatement (oracle, -O3 -march=native built on this machine), one pthread per METIS
    partition with shared-memory halo copies standing in for MPI, all host cores."""
    import torch
    from fesom2_b200 import fields as F, mesh as M, partition as P
    from oracle import oracle_py as O
    cores = os.cpu_count() or 1
    nthr = max(1, min(cores, args.cpu_threads or cores))
    torch.set_num_threads(cores)
    st = F.make_state(g, "cpu")
    dt = F.cfl_dt(g, st, 0.3)
    trs = F.make_tracers(g, ntr, "cpu", hor="MFCT", ver="QR4C", lim="FCT")
    part = P.partition(g, nthr, "metis") if nthr > 1 else np.zeros(g.Nh, np.int32)
    ranks = []
    for r in range(nthr):
        loc = M.localize(g, part, r) if nthr > 1 else g
        lst, ltr = F.scatter_to_local(g, loc, st, trs) if nthr > 1 else (st, trs)
        ranks.append(O.OracleRank(loc, lst, ltr, M.nboundary_lay(loc)))
    del st, trs
    O.build(fast=True)
    if warmup:
        O.run(ranks, dt, warmup, 0, fast=True)
    sec = O.run(ranks, dt, steps, 0, fast=True)
    units = float(g.N) * g.L * ntr
    return {"value": units * steps / sec, "unit": UNIT, "cores": nthr, "kind": "port",
            "sample": f"{steps} steps (+{warmup} warm-up) of {desc}, {ntr} tracers, {nthr} METIS partitions as threads",
            "ms_per_step": 1e3 * sec / steps}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg4-full", "core2", "recom30", "pi"])
    ap.add_argument("--tracers", type=int, default=0, help="tracers per batched call (default 2; recom30: 30)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--partitioner", default="metis", choices=["metis", "rcb"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ngpus = max(args.gpus, 1)
    ntr = args.tracers or (30 if args.workload == "recom30" else 2)
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = max(args.steps, 1)

    # -------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        g, desc = workload_mesh(args.workload, 1)      # bounded sample: the 1-GPU share of the workload
        steps = min(K, 5)
        res = cpu_arm(args, g, desc, steps, min(W, 1), ntr)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": ngpus,
                "steps": steps, "warmup": min(W, 1), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "tracers": ntr, "schemes": "MFCT+QR4C+FCT",
                           "note": "reference Fortran/MPI path cannot be built here (no gfortran/MPI); this is the C restatement (oracle)"},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # -------------------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from fesom2_b200 import fields as F, mesh as M, partition as P
    from fesom2_b200.driver import AdvB200, unique_id

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a GPU (no CPU fallback)")
    if world != ngpus:
        raise SystemExit(f"--gpus {ngpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {ngpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t_setup = time.time()
    g, desc = workload_mesh(args.workload, ngpus)
    if world > 1:
        part_t = torch.zeros(g.Nh, dtype=torch.int32, device=dev)
        if rank == 0:
            part_t.copy_(torch.from_numpy(P.partition(g, world, args.partitioner)))
        dist.broadcast(part_t, 0)
        loc = M.localize(g, part_t.cpu().numpy(), rank)
    else:
        loc = g
    gN, gL = g.N, g.L
    if world > 1:
        del g
    nb = M.nboundary_lay(loc)
    st = F.make_state(loc, dev)
    dt = F.cfl_dt(loc, st, 0.3)
    if world > 1:
        dtt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(dtt, op=dist.ReduceOp.MIN)
        dt = float(dtt.item())
    tri = F.find_up_downwind_triangles(loc)
    trs = []
    for k in range(ntr):
        trs += F.make_tracers_kind(loc, k, dev, tri, hor="MFCT", ver="QR4C", lim="FCT")
    dh = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in trs]
    ctx = AdvB200(loc, nb, device=local_rank, max_tracers=ntr)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
        # halos of the inputs (the reference's caller has exchanged values / valuesAB)
        ctx.exchange_nod([t.values for t in trs] + [t.valuesAB for t in trs], loc.L)
        ctx.synchronize()
    ctx.set_state(st)
    t_setup = time.time() - t_setup

    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ctx.synchronize()

    def step():
        # a model step hands over new uv/w/thicknesses before the tracer loop: the volume-flux
        # kernel is part of every timed step (device pointers: no copy, only Q is recomputed)
        ctx.set_state(st)
        ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False)

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    barrier()
    ev0.record(ext)
    for _ in range(K):
        step()
    ev1.record(ext)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / K
    units_step = float(gN) * gL * ntr                       # whole job, all ranks
    value = units_step / (ms_per_step * 1e-3)

    # ---- roofline of the step's kernels (events between kernels on the compute stream, N=1 only)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = PEAK_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    alg = algorithmic_bytes(loc.L, loc.N, loc.E, loc.T, ntr)
    roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "algorithmic_bytes_per_step_per_gpu": alg,
            "kernel": "whole step = k_edge_flux + k_node_lo + k_fct_bounds + k_fct_update (section-8d bytes are per step)"}
    if world == 1:
        ctx.set_profiling(True)
        ph = np.zeros(4)
        reps = 5
        for _ in range(reps):
            step()
            ctx.synchronize()
            ph += np.array(ctx.phase_ms()[:4])
        ctx.set_profiling(False)
        ph /= reps
        ksum = float(ph[0] + ph[1] + ph[2] + ph[3])
        roof["achieved"] = alg / (ksum * 1e-3) / 1e9
        roof["kernel_ms"] = {"k_edge_flux": float(ph[0]), "k_node_lo": float(ph[1]),
                             "k_fct_bounds": float(ph[2]), "k_fct_update": float(ph[3])}
    else:
        # per-rank algorithmic bytes over the max-over-ranks step time (includes exposed halo waits)
        roof["achieved"] = alg / (ms_per_step * 1e-3) / 1e9
    roof["frac"] = roof["achieved"] / peak
    # measured DRAM bytes of the same step from the committed `ncu --set full` capture
    # (tools/ncu_traffic.py -> profiles/traffic_latest.json; per launch, like `achieved`)
    traffic_file = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(traffic_file) and world == 1 and args.workload == "cfg4" and ntr == 2:
        try:
            tj = json.load(open(traffic_file))
            roof["traffic"] = tj.get("dram_bytes_per_step")
            roof["traffic_source"] = "profiles/traffic_latest.json (" + os.path.basename(tj.get("report", "ncu")) + ")"
            if "kernel_ms" in roof:
                roof["kernel_dram_gbs"] = {
                    k: (v["dram_read_bytes"] + v["dram_write_bytes"]) / (roof["kernel_ms"].get(k, float("nan")) * 1e-3) / 1e9
                    for k, v in tj.get("kernels", {}).items() if k in roof["kernel_ms"]}
        except Exception:
            pass

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        h_st = F.OceanState(**{k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in st.__dict__.items()})
        h_trs = [F.TracerFields(values=t.values.cpu().pin_memory(), valuesAB=t.valuesAB.cpu().pin_memory(),
                                edge_up_dn_grad=t.edge_up_dn_grad.cpu().pin_memory(), tra_adv_hor=t.tra_adv_hor,
                                tra_adv_ver=t.tra_adv_ver, tra_adv_lim=t.tra_adv_lim, tra_adv_ph=t.tra_adv_ph,
                                tra_adv_pv=t.tra_adv_pv) for t in trs]
        h_dh = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64).pin_memory() for _ in trs]
        h_dv = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64).pin_memory() for _ in trs]
        st_bytes = sum(v.numel() * 8 for k, v in h_st.__dict__.items() if torch.is_tensor(v) and k != "zbar_n_bot")
        h2d = st_bytes + sum(t.values.numel() * 8 * 4 + t.edge_up_dn_grad.numel() * 8 for t in h_trs)
        d2h = sum(x.numel() * 8 * 2 for x in h_dh)

        def e2e_step():
            ctx.set_state(h_st)                              # per-step `!$ACC UPDATE DEVICE` of the reference
            ctx.do_oce_adv_tra(dt, h_trs, h_dh, h_dv)        # blocking; copies tendencies back

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        # the same call with edge_up_dn_grad = NULL: the library computes the gradients on the device
        # (one rank) instead of receiving 4 E L words per tracer over PCIe -- reported
        # beside the contract number, not instead of it
        e2e_dg = None
        if world == 1:
            try:
                ctx.set_gradient_mesh(tri)
                g_keep = [t.edge_up_dn_grad for t in h_trs]
                for t in h_trs:
                    t.edge_up_dn_grad = None
                e2e_step()
                barrier()
                t0g = time.perf_counter()
                for _ in range(args.e2e_steps):
                    e2e_step()
                barrier()
                secg = time.perf_counter() - t0g
                e2e_dg = {"value": units_step * args.e2e_steps / secg, "unit": UNIT, "ms_per_step": 1e3 * secg / args.e2e_steps,
                          "h2d_bytes_per_step": int(h2d - sum(g.numel() * 8 for g in g_keep)), "d2h_bytes_per_step": int(d2h),
                          "note": "edge_up_dn_grad = NULL: tracer_gradient_elements + fill_up_dn_grad run on the device inside the call"}
                for t, gk in zip(h_trs, g_keep):
                    t.edge_up_dn_grad = gk
            except Exception as ex:
                e2e_dg = {"value": None, "note": f"failed: {ex}"}
        e2e = {"value": units_step * args.e2e_steps / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * sec / args.e2e_steps,
               "note": "adv_ctx_set_state + adv_do_oce_adv_tra with pinned HOST pointers; bytes are per rank"}
        if e2e_dg is not None:
            e2e["device_gradients"] = e2e_dg
        ctx.set_state(st)
        del h_st, h_trs, h_dh, h_dv

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpus, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.workload == "cfg4-full" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "tracers": ntr, "schemes": "MFCT+QR4C+FCT (ph=0, pv=1)", "global_nodes": int(gN),
                       "layers": int(gL), "nodes_per_rank": int(loc.N), "halo_nodes_rank0": int(loc.eDim_nod2D),
                       "partitioner": args.partitioner if world > 1 else "none", "dt_s": dt,
                       "l2": "inputs per step (>= 10 GB per GPU) exceed the 126 MB L2; no explicit flush",
                       "timing": "CUDA events on the library's compute stream, max over ranks", "setup_s": round(t_setup, 1)},
            "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del trs, dh, dv, st
        ctx.close()
        torch.cuda.empty_cache()
        try:
            res = cpu_arm(args, loc, desc, args.cpu_steps, 1, ntr)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

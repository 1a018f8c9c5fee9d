#!/usr/bin/env python
"""bench.py -- tracer node-level updates/s of the batched do_oce_adv_tra step on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      CPU restatement on the host cores

One "step" = one call of do_oce_adv_tra for the whole tracer batch (T+S, MFCT + QR4C + FCT),
including its two internal halo exchanges.  Default workload = BASELINE.json
config 4 as written: the 1733x1733x70 (3.0M-node) synthetic mesh at EVERY N, METIS-partitioned for
N > 1 -> strong scaling.  `--workload cfg4` is the weak-scaling variant of round 1 (same nodes per GPU).

`value`   : device-resident inputs, CUDA events on the library's compute stream, max over ranks.
`e2e`     : the same step through the C ABI with HOST buffers, H2D/D2H inside the timing.
`roofline`: SURVEY 8d algorithmic bytes of the step / summed kernel time, against MEASURED_PEAKS.
`parity`  : one step of the CPU oracle on the SAME mesh, partition and inputs (rank r's oracle exchanges
            halos with the other ranks' oracles over gloo), compared on every owned node.
"""
from __future__ import annotations

import argparse
import gc
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
SCHEMES = "MFCT+QR4C+FCT (ph=0, pv=1)"

# name -> (nx, ny, nl, description)
WORKLOADS = {
    "cfg4-full": (1733, 1733, 71, "BASELINE config 4: synthetic 1733x1733 (3.0M-node) lon-lat mesh, 70 layers, the whole mesh at every N (strong scaling)"),
    "cfg4": (None, None, 71, "BASELINE config 4, weak-scaling variant: per-GPU share of the 3.0M-node mesh (613x613 per GPU), 70 layers"),
    "core2": (357, 356, 48, "BASELINE config 3: synthetic CORE2-sized 357x356 mesh (127k nodes), 47 layers"),
    "recom30": (357, 356, 48, "BASELINE config 5: 30 tracers on the CORE2-sized 357x356 mesh, 47 layers"),
    "pi": (None, None, 48, "BASELINE config 1: pi test mesh (3140 nodes, 47 layers)"),
}


def workload_dims(name: str, ngpus: int):
    nx, ny, nl, desc = WORKLOADS[name]
    if name == "cfg4":
        nx = ny = int(round(1733 * math.sqrt(ngpus / 8.0)))
    return nx, ny, nl, desc


def workload_mesh(name: str, ngpus: int):
    from fesom2_b200 import mesh as M
    if name == "pi":
        return M.load_npz_mesh(os.path.join(ROOT, "tests", "golden", "mesh_pi.npz"))
    nx, ny, nl, _ = workload_dims(name, ngpus)
    return M.synth_mesh(nx, ny, nl=nl)


def config_of(name: str, ngpus: int, ntr: int) -> dict:
    """The `config` object: identical in the B200 arm and in the reference arm of the same command line."""
    nx, ny, nl, desc = workload_dims(name, ngpus)
    nodes = 3140 if name == "pi" else nx * ny
    return {"workload": f"{name}: {desc}", "tracers": ntr, "schemes": SCHEMES, "global_nodes": int(nodes), "layers": int(nl - 1),
            "partitioner": "metis (fort_part.c options, edge weights nlev(i)+nlev(j))" if ngpus > 1 else "none",
            "l2": "inputs per step exceed the 126 MB L2 many times over; no explicit flush",
            "timing": "CUDA events on the library's compute stream, max over ranks"}


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
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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


def cpu_sample_mesh(name: str):
    """Bounded sample of the workload for the CPU legs: the 1/8 share of config 4 (613x613x70, the per-GPU share of the
    8-GPU run), the mesh itself for the small configs."""
    from fesom2_b200 import mesh as M
    if name in ("cfg4-full", "cfg4"):
        return M.synth_mesh(613, 613, nl=71), "a 1/8 sample of the workload: synthetic 613x613 lon-lat mesh, 70 layers (one GPU's share of the 3.0M-node mesh)"
    return workload_mesh(name, 1), WORKLOADS[name][3]


# ------------------------------------------------------------------------------------------------
def gloo_exchange_fn(loc, group):
    """exchange_nod3D of one oracle rank over torch.distributed (gloo), following com_nod2D
    (src/gen_halo_exchange.F90:432-517): the received columns are the contiguous halo tail."""
    import torch
    import torch.distributed as dist
    com = loc.com_nod2D
    N = loc.N

    def fn(field: np.ndarray, nlev: int):
        t = torch.from_numpy(field)
        reqs, keep = [], []
        for i, p in enumerate(com.rPE):
            a, b = int(com.rptr[i]) - 1, int(com.rptr[i + 1]) - 1
            reqs.append(dist.irecv(t[N + a:N + b], src=int(p), group=group))
        for i, p in enumerate(com.sPE):
            seg = torch.as_tensor(com.slist[int(com.sptr[i]) - 1:int(com.sptr[i + 1]) - 1].astype(np.int64) - 1)
            buf = t[seg].contiguous()
            keep.append(buf)
            reqs.append(dist.isend(buf, dst=int(p), group=group))
        for r in reqs:
            r.wait()
    return fn


def parity_check(loc, nb, st, trs, dt, dh_gpu, dv_gpu, world, gloo_group):
    """One oracle step on this rank's mesh and inputs (outside every timed region), tracer by tracer, against the
    tendencies the library produced from zeroed del_ttf arrays.  Returns the rank-local result."""
    import torch
    from oracle import oracle_py as O
    t0 = time.time()
    snp = {k: v.cpu().numpy() for k, v in st.__dict__.items() if torch.is_tensor(v)}
    lo = O.LeanOracle(loc, snp, nb, use_wsplit=st.use_wsplit, fast=False)
    del snp
    xfn = gloo_exchange_fn(loc, gloo_group) if world > 1 else None
    N = loc.N
    max_rel, identical, ref_max = 0.0, True, 0.0
    for k, t in enumerate(trs):
        grad = t.edge_up_dn_grad.cpu().numpy() if t.edge_up_dn_grad is not None else None
        dh, dv = lo.run_tracer(t.values.cpu().numpy(), t.valuesAB.cpu().numpy(), grad, t.tra_adv_hor, t.tra_adv_ver,
                               t.tra_adv_lim, t.tra_adv_ph, t.tra_adv_pv, dt, exchange=xfn)
        del grad
        for got_t, ref in ((dh_gpu[k], dh), (dv_gpu[k], dv)):
            got = got_t[:N].cpu().numpy()
            ref = ref[:N]
            den = max(float(np.abs(ref).max()), 1e-300)
            max_rel = max(max_rel, float(np.abs(got - ref).max()) / den)
            identical = identical and bool(np.array_equal(got, ref))
            ref_max = max(ref_max, den)
        del dh, dv
    return {"max_rel_err": max_rel, "bit_identical": identical, "oracle_s": time.time() - t0, "ref_max": ref_max}


# ------------------------------------------------------------------------------------------------
class Case:
    """One workload resident on this rank's GPU: mesh, state, tracers, context."""
    _host_keep = None

    def __init__(self, name, ngpus, ntr, rank, world, local_rank, partitioner, dist):
        import torch
        from fesom2_b200 import fields as F, mesh as M, partition as P
        from fesom2_b200.driver import AdvB200, unique_id
        self.name, self.ntr, self.world, self.rank = name, ntr, world, rank
        dev = self.dev = torch.device(f"cuda:{local_rank}")
        t0 = time.time()
        g = workload_mesh(name, ngpus)
        self.gN, self.gL = g.N, g.L
        if world > 1:
            part_t = torch.zeros(g.Nh, dtype=torch.int32, device=dev)
            if rank == 0:
                part_t.copy_(torch.from_numpy(P.partition(g, world, partitioner)))
            dist.broadcast(part_t, 0)
            self.part = part_t.cpu().numpy()
            loc = M.localize(g, self.part, rank)
        else:
            self.part = None
            loc = g
        self.g = g
        self.loc = loc
        self.nb = M.nboundary_lay(loc)
        self.st = F.make_state(loc, dev)
        dt = F.cfl_dt(loc, self.st, 0.3)
        if world > 1:
            dtt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(dtt, op=dist.ReduceOp.MIN)
            dt = float(dtt.item())
        self.dt = dt
        self.ctx = ctx = AdvB200(loc, self.nb, device=local_rank, max_tracers=ntr)
        # inputs: T, S (+ RECOM-style copies); edge_up_dn_grad from the library's own producer on this rank's mesh
        # (tracer_gradient_elements + fill_up_dn_grad; memory-lean at 3.0M nodes) -- an input like any other: the
        # timed call receives it as the reference's do_oce_adv_tra does
        self.tri = F.find_up_downwind_triangles(loc, dev)
        ctx.set_gradient_mesh(self.tri)
        self.trs = []
        xy = torch.zeros((loc.T, loc.L, 2), dtype=torch.float64, device=dev)
        for k in range(ntr):
            v, vo = F.make_tracer_values(loc, dev, kind=k)
            vab = F.ab2(v, vo).contiguous()
            del vo
            grad = torch.zeros((loc.E, loc.L, 4), dtype=torch.float64, device=dev)
            ctx.tracer_gradient_elements([v], [xy])
            ctx.fill_up_dn_grad([xy], [grad])
            self.trs.append(F.TracerFields(values=v, valuesAB=vab, edge_up_dn_grad=grad, tra_adv_hor="MFCT", tra_adv_ver="QR4C",
                                           tra_adv_lim="FCT", tra_adv_ph=0.0, tra_adv_pv=1.0))
        ctx.synchronize()
        del xy
        self.dh = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in self.trs]
        self.dv = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in self.trs]
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
            # halos of the inputs (the reference's caller has exchanged values / valuesAB)
            ctx.exchange_nod([t.values for t in self.trs] + [t.valuesAB for t in self.trs], loc.L)
            ctx.synchronize()
        ctx.set_state(self.st)
        self.setup_s = time.time() - t0
        self.dist = dist

    def barrier(self):
        import torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)
        self.ctx.synchronize()

    def step(self):
        # a model step hands over new uv/w/thicknesses before the tracer loop: the volume-flux
        # computation is part of every timed step (device pointers: no copy, only Q is recomputed)
        self.ctx.set_state(self.st)
        self.ctx.do_oce_adv_tra(self.dt, self.trs, self.dh, self.dv, sync=False)

    def timed(self, K, W, sample_clocks):
        """W warm-up steps, then exactly K steps between barriers; device time, max over ranks."""
        import torch
        ctx = self.ctx
        ext = torch.cuda.ExternalStream(ctx.stream, device=self.dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(W):
            self.step()
        self.barrier()
        sampler = ClockSampler(self.dev.index) if sample_clocks else None
        launches0 = ctx.launch_count
        self.barrier()
        ev0.record(ext)
        for _ in range(K):
            self.step()
        ev1.record(ext)
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = ctx.launch_count - launches0
        clocks = sampler.stop() if sampler else None
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / K, launches, clocks

    def roofline(self, ms_per_step, peak, peak_src):
        loc = self.loc
        alg = algorithmic_bytes(loc.L, loc.N, loc.E, loc.T, self.ntr)
        roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                "algorithmic_bytes_per_step_per_gpu": alg,
                "kernel": "whole step = k_edge_flux + k_node_lo + k_fct_bounds + k_fct_update (section-8d bytes are per step)"}
        if self.world == 1:
            ctx = self.ctx
            ctx.set_profiling(True)
            ph = np.zeros(4)
            reps = 5
            for _ in range(reps):
                self.step()
                ctx.synchronize()
                ph += np.array(ctx.phase_ms()[:4])
            ctx.set_profiling(False)
            ph /= reps
            ksum = float(ph.sum())
            roof["achieved"] = alg / (ksum * 1e-3) / 1e9
            roof["kernel_ms"] = {"k_edge_flux": float(ph[0]), "k_node_lo": float(ph[1]),
                                 "k_fct_bounds": float(ph[2]), "k_fct_update": float(ph[3])}
        else:
            # per-rank algorithmic bytes over the max-over-ranks step time (includes exposed halo waits)
            roof["achieved"] = alg / (ms_per_step * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        return roof

    def halo(self):
        """exchange volume and overlap of one step (events on the communication and compute streams)"""
        import torch
        if self.world == 1:
            return None
        self.step()
        self.ctx.synchronize()
        b, comm, exposed = self.ctx.halo_stats()
        t = torch.tensor([float(b), comm[0] + comm[1], exposed[0] + exposed[1]], dtype=torch.float64, device=self.dev)
        tmax = t.clone()
        self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return {"bytes_per_step_all_ranks": int(t[0].item()), "bytes_per_step_max_rank": int(tmax[0].item()),
                "halo_nodes_rank0": int(self.loc.eDim_nod2D), "exchanges_per_step": 2,
                "comm_stream_ms_max_rank": float(tmax[1].item()), "exposed_ms_max_rank": float(tmax[2].item()),
                "note": "comm = pack + NCCL send/recv of both exchanges on the communication stream; exposed = time the compute "
                        "stream waited for them after the interior kernels (0 = fully hidden)"}

    def parity(self, gloo_group):
        import torch
        for x in self.dh + self.dv:
            x.zero_()
        torch.cuda.synchronize(self.dev)
        self.ctx.set_state(self.st)
        self.ctx.do_oce_adv_tra(self.dt, self.trs, self.dh, self.dv)
        res = parity_check(self.loc, self.nb, self.st, self.trs, self.dt, self.dh, self.dv, self.world, gloo_group)
        if self.world > 1:
            t = torch.tensor([res["max_rel_err"], 0.0 if res["bit_identical"] else 1.0, res["oracle_s"]], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            res = {"max_rel_err": float(t[0].item()), "bit_identical": bool(t[1].item() == 0.0), "oracle_s": float(t[2].item()), "ref_max": res["ref_max"]}
        res["checked"] = ("every owned node of every rank, all tracers, del_ttf_advhoriz and del_ttf_advvert of one step from zeroed "
                          "tendencies; oracle = C restatement, one rank per partition" + (", halos exchanged over gloo" if self.world > 1 else ""))
        res["tolerance"] = 1e-12
        res["ok"] = bool(res["max_rel_err"] <= 1e-12)
        return res

    def e2e(self, steps):
        """The same step through the C ABI with HOST buffers: plain pageable arrays, as a Fortran host's allocatables
        are; the library page-locks them on first use (cudaHostRegister, inside the warm-up call)."""
        import torch
        from fesom2_b200 import fields as F
        ctx, loc, st, trs = self.ctx, self.loc, self.st, self.trs
        h_st = F.OceanState(**{k: (v.cpu() if torch.is_tensor(v) else v) for k, v in st.__dict__.items()})
        h_trs = [F.TracerFields(values=t.values.cpu(), valuesAB=t.valuesAB.cpu(), edge_up_dn_grad=t.edge_up_dn_grad.cpu(),
                                tra_adv_hor=t.tra_adv_hor, tra_adv_ver=t.tra_adv_ver, tra_adv_lim=t.tra_adv_lim,
                                tra_adv_ph=t.tra_adv_ph, tra_adv_pv=t.tra_adv_pv) for t in trs]
        h_dh = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64) for _ in trs]
        h_dv = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64) for _ in trs]
        # the host path stages its own device copies: release the device-resident inputs of the other legs first
        # (3.0M nodes x 70 layers: 76 GB of inputs; e2e is the last leg that uses this case)
        ctx.set_state(h_st)
        ctx.synchronize()
        self._host_keep = (h_st, h_trs, h_dh, h_dv)        # page-locked by the library: must outlive the context
        for t in trs:
            t.values = t.valuesAB = t.edge_up_dn_grad = None
        self.st, self.trs, self.dh, self.dv = None, [], [], []
        del st, trs
        gc.collect()
        torch.cuda.empty_cache()
        st_bytes = sum(v.numel() * 8 for k, v in h_st.__dict__.items() if torch.is_tensor(v) and k != "zbar_n_bot")
        h2d = st_bytes + sum(t.values.numel() * 8 * 4 + t.edge_up_dn_grad.numel() * 8 for t in h_trs)
        d2h = sum(x.numel() * 8 * 2 for x in h_dh)
        units_step = float(self.gN) * self.gL * self.ntr

        def e2e_step():
            ctx.set_state(h_st)                              # per-step `!$ACC UPDATE DEVICE` of the reference
            ctx.do_oce_adv_tra(self.dt, h_trs, h_dh, h_dv)   # blocking; copies tendencies back

        def run():
            e2e_step()                                       # warm-up: staging buffers, page-locking
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            self.barrier()
            sec = time.perf_counter() - t0
            if self.world > 1:
                t = torch.tensor([sec], dtype=torch.float64, device=self.dev)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                sec = float(t.item())
            return sec

        sec = run()
        out = {"value": units_step * steps / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * sec / steps, "steps": steps,
               "note": "adv_ctx_set_state + adv_do_oce_adv_tra with pageable HOST arrays (page-locked by the library on first use); bytes are per rank"}
        # the same call with edge_up_dn_grad = NULL: the library computes the gradients on the device (SURVEY 8f row 1;
        # exchange_elem over NCCL on N ranks) instead of receiving 4 E L words per tracer over PCIe
        try:
            if self.world > 1:
                from fesom2_b200 import fields as F2, mesh as M
                tri_g = F2.find_up_downwind_triangles(self.g, self.dev)
                ctx.set_gradient_mesh(gmesh=M.gradient_mesh(self.g, tri_g, self.part, loc))
            g_keep = [t.edge_up_dn_grad for t in h_trs]
            for t in h_trs:
                t.edge_up_dn_grad = None
            secg = run()
            out["device_gradients"] = {"value": units_step * steps / secg, "unit": UNIT, "ms_per_step": 1e3 * secg / steps,
                                       "h2d_bytes_per_step": int(h2d - sum(x.numel() * 8 for x in g_keep)), "d2h_bytes_per_step": int(d2h),
                                       "note": "edge_up_dn_grad = NULL: tracer_gradient_elements (+ exchange_elem) + fill_up_dn_grad run on the device inside the call"}
            del g_keep
        except Exception as ex:
            out["device_gradients"] = {"value": None, "note": f"failed: {ex}"}
        ctx.synchronize()
        return out

    def close(self):
        import torch
        self.ctx.close()
        self.trs = self.dh = self.dv = self.st = self._host_keep = None
        gc.collect()
        torch.cuda.empty_cache()


def sub_bench(name, ntr, steps, peak, peak_src, dist):
    """A secondary configuration on one GPU (embedded in the N=1 line): ms, roofline fraction, parity."""
    c = Case(name, 1, ntr, 0, 1, 0, "metis", dist)
    ms, launches, _ = c.timed(steps, 3, False)
    roof = c.roofline(ms, peak, peak_src)
    par = c.parity(None)
    out = {"workload": f"{name}: {WORKLOADS[name][3]}", "tracers": ntr, "steps": steps, "ms_per_step": ms,
           "value": float(c.gN) * c.gL * ntr / (ms * 1e-3), "unit": UNIT, "roofline_frac": roof["frac"],
           "kernel_ms": roof.get("kernel_ms"), "gpu_launches": int(launches),
           "parity": {k: par[k] for k in ("max_rel_err", "bit_identical", "ok")}}
    c.close()
    return out


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank: int):
    """What `srun --cpu-bind` / `numactl` do for the one-rank-per-GPU launch of an MPI host (the reference's
    work/job_gpu_levante binds ranks next to their GPUs): pin this rank to the CPUs of its GPU's NUMA node, so that the
    first-touch pages of its host arrays sit behind the GPU's own PCIe root (the e2e leg copies 9.5 GB per rank and step).
    Returns the number of CPUs bound to, or 0 when the affinity is unknown (then nothing changes)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode() if hasattr(bus, "encode") else bus)
        words = (max(os.cpu_count() or 64, 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4-full", choices=sorted(WORKLOADS))
    ap.add_argument("--tracers", type=int, default=0, help="tracers per batched call (default 2; recom30: 30)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the embedded secondary configurations (N=1)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--partitioner", default="metis", choices=["metis", "rcb"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ngpus = max(args.gpus, 1)
    ntr = args.tracers or (30 if args.workload == "recom30" else 2)
    W = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    K = max(args.steps, 1)
    config = config_of(args.workload, ngpus, ntr)

    # -------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        g, sdesc = cpu_sample_mesh(args.workload)
        res = cpu_arm(args, g, sdesc, K, W, ntr)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": ngpus,
                "steps": K, "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak" if args.workload == "cfg4" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "note": "the reference's Fortran/MPI path cannot be built here or on the GPU box (no Fortran compiler, no MPI: "
                        "gpurun_out/r4a_probe.log); this is its C restatement (oracle), one thread per METIS partition; each step "
                        "runs the bounded sample named in cpu_baseline.sample, value = sample units / sample time",
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # -------------------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a GPU (no CPU fallback)")
    if world != ngpus:
        raise SystemExit(f"--gpus {ngpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {ngpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    gloo_group = None
    if world > 1:
        ncpu = bind_to_gpu_numa(local_rank)
        config["cpu_bind"] = f"each rank pinned to the CPUs of its GPU's NUMA node ({ncpu} on rank 0)" if ncpu else "none (GPU affinity unknown)"
        dist.init_process_group("nccl", device_id=dev)
        if not args.no_parity:
            gloo_group = dist.new_group(backend="gloo")

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = PEAK_FALLBACK_GBS, "fallback (B200_PROFILING.md)"

    case = Case(args.workload, ngpus, ntr, rank, world, local_rank, args.partitioner, dist)
    loc = case.loc
    ms_per_step, launches, clocks = case.timed(K, W, rank == 0)
    units_step = float(case.gN) * case.gL * ntr                       # whole job, all ranks
    value = units_step / (ms_per_step * 1e-3)
    roof = case.roofline(ms_per_step, peak, peak_src)
    # measured DRAM bytes of the step from the committed `ncu --set full` capture of the per-GPU share (613x613x70)
    # (tools/ncu_traffic.py -> profiles/traffic_latest.json; per launch, like `achieved`), scaled by the node count
    traffic_file = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(traffic_file) and world == 1 and args.workload in ("cfg4", "cfg4-full") and ntr == 2:
        try:
            tj = json.load(open(traffic_file))
            scale = float(loc.N) / float(tj.get("nodes", 375769))
            roof["traffic"] = tj.get("dram_bytes_per_step") * scale
            roof["traffic_source"] = ("profiles/traffic_latest.json (" + os.path.basename(tj.get("report", "ncu")) +
                                      f", 613x613x70 capture x {scale:.3f} = node ratio)")
        except Exception:
            pass
    def leg(name, fn):
        t0 = time.time()
        try:
            out = fn()
        except Exception as ex:   # one failing leg must not take the whole line down
            out = {"failed": f"{type(ex).__name__}: {ex}"}
        if rank == 0:
            print(f"[bench] {name}: {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
        return out

    if rank == 0:
        print(f"[bench] setup {case.setup_s:.1f} s, {ms_per_step:.3f} ms/step", file=sys.stderr, flush=True)
    halo = leg("halo", case.halo)
    parity = None if args.no_parity else leg("parity", lambda: case.parity(gloo_group))
    e2e = None if args.no_e2e else leg("e2e", lambda: case.e2e(args.e2e_steps))
    if e2e is not None and "failed" in e2e:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "note": "failed: " + e2e["failed"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpus, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.workload == "cfg4" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "run": {"nodes_per_rank": int(loc.N), "halo_nodes_rank0": int(loc.eDim_nod2D), "dt_s": case.dt, "setup_s": round(case.setup_s, 1)},
            "roofline": roof, "parity": parity, "halo": halo, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    case.close()
    del case

    if rank == 0 and world == 1:
        if not args.no_sub and args.workload == "cfg4-full":
            subs = {}
            for nme, nt, st_ in (("cfg4", 2, 10), ("core2", 2, 20), ("recom30", 30, 5)):
                try:
                    subs[nme] = sub_bench(nme, nt, st_, peak, peak_src, dist)
                except Exception as ex:
                    subs[nme] = {"failed": str(ex)}
            line["other_configs"] = subs
        if not args.no_cpu_baseline:
            try:
                g, sdesc = cpu_sample_mesh(args.workload)
                res = cpu_arm(args, g, sdesc, args.cpu_steps, 1, ntr)
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

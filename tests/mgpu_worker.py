"""Multi-rank worker used by the world_size>1 tests (one process per rank).

backend nccl : each rank drives one GPU through the C ABI (adv_ctx_comm_init + the library's own
               packed-halo NCCL exchange) and returns its owned-node tendencies.
backend gloo : CPU-only check of the HOST logic of the N>1 path: partition -> local numbering ->
               com_nod2D send/recv lists; the exchange itself is done with torch.distributed
               send/recv following exactly those lists.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fesom2_b200 import fields as F  # noqa: E402
from fesom2_b200 import mesh as M  # noqa: E402


def build_global(mesh_name: str, npes: int, hor="MFCT", ver="QR4C", ntr=2):
    if mesh_name == "synth":
        g = M.synth_mesh(41, 37, nl=24, min_layers=4)
        from fesom2_b200 import partition as P
        part = P.partition(g, npes, "metis")
    else:
        g = M.load_npz_mesh(os.path.join(ROOT, "tests", "golden", f"mesh_{mesh_name}.npz"))
        part = g.parts[npes]                      # the reference's checked-in dist_N partition
    st = F.make_state(g, "cpu")
    dt = F.cfl_dt(g, st, 0.3)
    trs = F.make_tracers(g, ntr, "cpu", hor=hor, ver=ver, lim="FCT")
    return g, part, st, trs, dt


def gloo_exchange(loc: M.Mesh, field: torch.Tensor):
    """exchange_nod over torch.distributed following com_nod2D (gen_halo_exchange.F90:432-517)."""
    com = loc.com_nod2D
    reqs, bufs = [], []
    for i, p in enumerate(com.rPE):
        seg = com.rlist[com.rptr[i] - 1:com.rptr[i + 1] - 1] - 1
        buf = torch.empty((len(seg), field.shape[1]), dtype=field.dtype)
        bufs.append((seg, buf))
        reqs.append(dist.irecv(buf, src=int(p)))
    for i, p in enumerate(com.sPE):
        seg = com.slist[com.sptr[i] - 1:com.sptr[i + 1] - 1] - 1
        reqs.append(dist.isend(field[torch.as_tensor(seg.astype(np.int64))].contiguous(), dst=int(p)))
    for r in reqs:
        r.wait()
    for seg, buf in bufs:
        field[torch.as_tensor(seg.astype(np.int64))] = buf


def worker(rank: int, world: int, backend: str, mesh_name: str, port: int, out_dir: str, nsteps: int = 1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend == "nccl":
        torch.cuda.set_device(rank)
    dist.init_process_group(backend, rank=rank, world_size=world)
    g, part, st, trs, dt = build_global(mesh_name, world)
    loc = M.localize(g, part, rank)
    lst, ltr = F.scatter_to_local(g, loc, st, trs)
    nb = M.nboundary_lay(g)[loc.myList_nod2D - 1]      # global value (the reference exchanges it implicitly via edges)
    res = {}
    if backend == "gloo":
        # halo consistency: owners' values must arrive in the halo tail
        gid = torch.as_tensor(loc.myList_nod2D.astype(np.float64))
        fld = (gid[:, None] * 1000.0 + torch.arange(loc.L, dtype=torch.float64)[None, :]).contiguous()
        ref = fld.clone()
        fld[loc.N:] = -1.0
        gloo_exchange(loc, fld)
        res["halo_ok"] = bool(torch.equal(fld, ref))
        res["N"] = loc.N
    else:
        from common import to_device
        from fesom2_b200.driver import AdvB200, unique_id
        dev = torch.device(f"cuda:{rank}")
        ctx = AdvB200(loc, nb, device=rank, max_tracers=len(ltr))
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
        st_d, trs_d = to_device(lst, ltr, dev)
        dh = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in ltr]
        dv = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in ltr]
        ctx.set_state(st_d)
        for step in range(nsteps):
            if step > 0:
                for x in dh + dv:
                    x.zero_()
            ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
            if nsteps > 1:
                ctx.update_values([t.values for t in trs_d], dh, dv)   # + exchange_nod(values)
                ctx.synchronize()
        res["dh"] = [x.cpu().numpy() for x in dh]
        res["dv"] = [x.cpu().numpy() for x in dv]
        res["values"] = [t.values.cpu().numpy() for t in trs_d]
        res["launches"] = ctx.launch_count
        ctx.close()
    res["owned"] = loc.myList_nod2D[:loc.N]
    res["all_nodes"] = loc.myList_nod2D
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":   # torchrun entry: backend mesh out_dir [nsteps]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    worker(rank, world, sys.argv[1], sys.argv[2], int(os.environ.get("MASTER_PORT", "29533")), sys.argv[3],
           int(sys.argv[4]) if len(sys.argv) > 4 else 1)

#!/usr/bin/env python
"""tools/traffic_model.py [--side 613] [--nl 71] [--tracers 2] [traffic.json]

Expected DRAM bytes of each of the four step kernels on the bench mesh, counted from the mesh itself (wet
layers only; every array touched once per kernel, gathers served by L2), beside the ncu measurement in
profiles/traffic_latest.json and the contract's algorithmic bytes.  CPU only."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=613)
    ap.add_argument("--nl", type=int, default=71)
    ap.add_argument("--tracers", type=int, default=2)
    ap.add_argument("traffic", nargs="?", default=os.path.join(ROOT, "profiles", "traffic_latest.json"))
    a = ap.parse_args()
    from fesom2_b200 import mesh as M
    g = M.synth_mesh(a.side, a.side, nl=a.nl)
    B, L = a.tracers, g.L
    wn = (np.asarray(g.nlevels_nod2D) - np.asarray(g.ulevels_nod2D)).astype(np.int64)      # wet layers per node
    we = (np.asarray(g.nlevels) - np.asarray(g.ulevels)).astype(np.int64)                  # per element
    et = np.asarray(g.edge_tri).astype(np.int64) - 1
    hi = np.maximum(we[et[:, 0]], np.where(et[:, 1] >= 0, we[np.maximum(et[:, 1], 0)], 0))  # scatter range per edge
    WN, WT, WE = int(wn.sum()), int(we.sum()), int(hi.sum())
    w8 = 8.0
    model = {
        # edge kernel: grad 4 B, uv 2 + helem 1 per element layer, ttf + ttfAB per node layer; writes adf_h B + Q
        "k_edge_flux_b": w8 * (4 * B * WE + 3 * WT + 2 * B * WN + B * WE + WE),
        # N1: Q per edge layer, ttf, ttfAB (B each), 8 geometry/state words per node layer; writes lo B, adf_v B
        "k_node_lo": w8 * (WE + 2 * B * WN + 8 * WN + 2 * B * WN),
        # K2: lo B, ttf B, adf_h B per edge layer, adf_v B, areasvol, hnode_new; writes R+- 2B
        "k_fct_bounds": w8 * (2 * B * WN + B * WE + B * WN + 2 * WN + 2 * B * WN),
        # K3: adf_h B per edge layer, adf_v B, R+- 2B, lo B, ttf B, del_ttf 2B read + 2B write, 3 geometry words
        "k_fct_update": w8 * (B * WE + B * WN + 2 * B * WN + 2 * B * WN + 4 * B * WN + 3 * WN),
    }
    alg = w8 * L * (B * (12.0 * g.N + 4.0 * g.E) + (3.0 * g.T + 7.0 * g.N))
    meas = json.load(open(a.traffic)) if os.path.exists(a.traffic) else {"kernels": {}}
    print(f"mesh {a.side}x{a.side}x{L}: N={g.N} T={g.T} E={g.E}; wet node-layers {WN} ({WN / (g.N * L):.3f}), "
          f"element-layers {WT} ({WT / (g.T * L):.3f}), edge-layers {WE} ({WE / (g.E * L):.3f})")
    print(f"{'kernel':16s} {'model GB':>9s} {'ncu GB':>8s} {'ncu/model':>9s}")
    tot_m = tot_n = 0.0
    for k, v in model.items():
        mk = meas["kernels"].get(k)
        n = (mk["dram_read_bytes"] + mk["dram_write_bytes"]) if mk else float("nan")
        tot_m += v; tot_n += n
        print(f"{k:16s} {v / 1e9:9.2f} {n / 1e9:8.2f} {n / v:9.2f}")
    print(f"{'step':16s} {tot_m / 1e9:9.2f} {tot_n / 1e9:8.2f} {tot_n / tot_m:9.2f}")
    print(f"contract algorithmic bytes (all {L} layers, no intermediates): {alg / 1e9:.2f} GB; "
          f"four-pass model / contract = {tot_m / alg:.2f}")


if __name__ == "__main__":
    main()

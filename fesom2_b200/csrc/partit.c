/* fesom2_b200/csrc/partit.c -- mesh partitioning for the multi-GPU path.
 *
 * Restates the single-level branch of the reference's `do_partit` (src/fort_part.c:46-241, built
 * with METIS_VERSION=5, PART_WEIGHTED, METISRANDOMSEED=35243: mesh_part/CMakeLists.txt:80-82):
 * METIS_PartGraphRecursive on the node graph with two balance constraints (2-D node count and
 * nlevels+100), edge weights nlev(i)+nlev(j) (USE_EDGE_WEIGHTS is defined unconditionally, fort_part.c:11,
 * :191-205), NCUTS=10, NITER=15, UFACTOR=1, Fortran numbering.  Links the METIS 5 that ships
 * with the CUDA toolkit (libmetis_static.a, 64-bit idx_t) instead of the vendored lib/metis-5.1.0.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef int64_t idx_t; /* width of the toolkit's libmetis_static.a (probed) */
int METIS_SetDefaultOptions(idx_t *options);
int METIS_PartGraphRecursive(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt,
                             idx_t *vsize, idx_t *adjwgt, idx_t *nparts, float *tpwgts, float *ubvec,
                             idx_t *options, idx_t *edgecut, idx_t *part);

/* METIS 5.1.0 option indices (metis.h moptions_et) */
enum { OPT_OBJTYPE = 1, OPT_NITER = 6, OPT_NCUTS = 7, OPT_SEED = 8, OPT_CONTIG = 11, OPT_UFACTOR = 16,
       OPT_NUMBERING = 17, NOPTIONS = 40 };

/* n nodes; ptr(n+1), adj: CSR node graph with 1-based values (ssh_stiff%rowptr/colind of
 * src/fvom_init.F90:1792); wgt(n): nlevels_nod2D; part(n) out, 0-based ranks.  Returns the edge
 * cut, or -1 on error. */
long long fesom_partit(int n, const int32_t *ptr, const int32_t *adj, const int32_t *wgt, int np, int32_t *part)
{
    if (np < 1) return -1;
    if (np == 1) { for (int i = 0; i < n; ++i) part[i] = 0; return 0; }  /* fort_part.c:66 */
    idx_t opt[NOPTIONS];
    METIS_SetDefaultOptions(opt);
    opt[OPT_CONTIG] = 0;        /* fort_part.c:103 */
    opt[OPT_OBJTYPE] = 0;       /* METIS_OBJTYPE_CUT, :107 */
    opt[OPT_NUMBERING] = 1;     /* :116 */
    opt[OPT_NCUTS] = 10;        /* :117 */
    opt[OPT_NITER] = 15;        /* :118 */
    opt[OPT_UFACTOR] = 1;       /* :120 */
    opt[OPT_SEED] = 35243;      /* :127 */
    idx_t nn = n, ncon = 2, npp = np, ec = 0;
    const size_t nnz = (size_t)ptr[n] - 1;
    idx_t *xadj = malloc(sizeof(idx_t) * ((size_t)n + 1));
    idx_t *adjn = malloc(sizeof(idx_t) * (nnz ? nnz : 1));
    idx_t *vw = malloc(sizeof(idx_t) * 2 * (size_t)n);
    idx_t *p = malloc(sizeof(idx_t) * (size_t)n);
    idx_t *ew = malloc(sizeof(idx_t) * (nnz ? nnz : 1));
    if (!xadj || !adjn || !vw || !p || !ew) return -1;
    for (int i = 0; i <= n; ++i) xadj[i] = ptr[i];
    for (size_t k = 0; k < nnz; ++k) adjn[k] = adj[k];
    for (int i = 0; i < n; ++i) { vw[2 * i] = 1; vw[2 * i + 1] = wgt ? wgt[i] + 100 : 100; }  /* :176-179 */
    for (int i = 0; i < n; ++i)                                                                /* :202-204 */
        for (int j = ptr[i] - 1; j < ptr[i + 1] - 1; ++j) ew[j] = wgt ? wgt[i] + wgt[adj[j] - 1] : 1;
    int rc = METIS_PartGraphRecursive(&nn, &ncon, xadj, adjn, vw, NULL, ew, &npp, NULL, NULL, opt, &ec, p); /* :233 */
    if (rc == 1) for (int i = 0; i < n; ++i) part[i] = (int32_t)(p[i] - 1);   /* :240 */
    free(xadj); free(adjn); free(vw); free(p); free(ew);
    return rc == 1 ? (long long)ec : -1;
}

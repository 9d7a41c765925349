/* TEST INFRASTRUCTURE ONLY (oracle/): CPU statement of the deterministic simplex
 * that solves the reference's fluid model.
 *
 * Reference: environments/class_FJSP.py:256-290 and class_MODFJSP.py:251-285
 * (`fluid_model`).  The reference hands this LP to IBM CPLEX via `docplex`, a
 * closed third-party dependency that is absent from /root/reference and from
 * this image; the LP optimum (max-min rate t*) is unique but the optimal VERTEX
 * that CPLEX returns is not, and the environment's later behaviour depends on
 * the vertex (which machines carry non-zero fluid rate).  The framework therefore
 * pins the vertex with its own fully specified pivoting rule ("fluid LP
 * specification"); this file is that specification in scalar C, the CUDA kernel
 * (csrc/fjsp_lp.cuh) performs the identical floating-point operations, and
 * tests/test_lp.py checks the optimum t* against scipy's HiGHS.
 *
 * Problem:  maximise z[t_col]  s.t.  A z <= b (b >= 0),  z >= 0.
 * Method :  revised primal simplex with an explicit dense basis inverse.
 *   start    : all-slack basis (feasible because b >= 0)
 *   pricing  : Dantzig (most negative reduced cost, ties -> lowest column) for
 *              the first DANTZIG_FACTOR*nrow+100 iterations, then Bland
 *   ratio    : min max(xB_i,0)/w_i over w_i > EPS_PIV, ties -> lowest basic var
 *   update   : explicit inverse, row p scaled first, every mul and add rounded
 *              separately (no FMA) and in the order written below
 *   output   : structural values below 1e-9 are reported as exactly 0
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPS_D 1e-9
#define EPS_PIV 1e-9
#define EPS_ZERO 1e-9
#define DANTZIG_FACTOR 20
#define HARD_FACTOR 200

int fjsp_lp_solve_sparse(int ncol, int nrow, const int *colptr, const int *rowidx,
                         const double *vals, const double *b, int t_col,
                         double *z, int *iters_out)
{
    int nvar = ncol + nrow;
    double *Binv = (double *)calloc((size_t)nrow * nrow, sizeof(double));
    double *xB = (double *)malloc(sizeof(double) * nrow);
    double *y = (double *)malloc(sizeof(double) * nrow);
    double *w = (double *)malloc(sizeof(double) * nrow);
    int *basis = (int *)malloc(sizeof(int) * nrow);
    int *pos = (int *)malloc(sizeof(int) * nvar);
    int rc = 0, it = 0;
    for (int j = 0; j < nvar; ++j) pos[j] = -1;
    for (int i = 0; i < nrow; ++i) {
        Binv[(size_t)i * nrow + i] = 1.0;
        xB[i] = b[i];
        basis[i] = ncol + i;
        pos[ncol + i] = i;
    }
    int dantzig_iters = DANTZIG_FACTOR * nrow + 100;
    int hard_iters = HARD_FACTOR * nrow + 1000;
    for (;; ++it) {
        if (it >= hard_iters) { rc = 2; break; }
        /* y = c_B^T Binv with c = -e_t (minimise -t) */
        int pt = pos[t_col];
        for (int k = 0; k < nrow; ++k) y[k] = (pt >= 0) ? -Binv[(size_t)pt * nrow + k] : 0.0;
        /* pricing */
        int q = -1;
        double dq = -EPS_D;
        int bland = it >= dantzig_iters;
        for (int j = 0; j < nvar; ++j) {
            if (pos[j] >= 0) continue;
            double d;
            if (j < ncol) {
                double acc = 0.0;
                for (int k = colptr[j]; k < colptr[j + 1]; ++k) acc = acc + y[rowidx[k]] * vals[k];
                d = ((j == t_col) ? -1.0 : 0.0) - acc;
            } else {
                d = -y[j - ncol];
            }
            if (d < dq) {
                q = j; dq = d;
                if (bland) break;
            }
        }
        if (q < 0) break; /* optimal */
        /* w = Binv * A_q */
        for (int i = 0; i < nrow; ++i) {
            if (q < ncol) {
                double acc = 0.0;
                for (int k = colptr[q]; k < colptr[q + 1]; ++k)
                    acc = acc + Binv[(size_t)i * nrow + rowidx[k]] * vals[k];
                w[i] = acc;
            } else {
                w[i] = Binv[(size_t)i * nrow + (q - ncol)];
            }
        }
        /* ratio test */
        int p = -1;
        double best = 0.0;
        for (int i = 0; i < nrow; ++i) {
            if (w[i] > EPS_PIV) {
                double xb = xB[i] > 0.0 ? xB[i] : 0.0;
                double r = xb / w[i];
                if (p < 0 || r < best || (r == best && basis[i] < basis[p])) { p = i; best = r; }
            }
        }
        if (p < 0) { rc = 3; break; } /* unbounded: cannot happen for the fluid model */
        /* update */
        double theta = best, wp = w[p];
        for (int i = 0; i < nrow; ++i)
            if (i != p) xB[i] = xB[i] - theta * w[i];
        xB[p] = theta;
        double *rowp = Binv + (size_t)p * nrow;
        for (int k = 0; k < nrow; ++k) rowp[k] = rowp[k] / wp;
        for (int i = 0; i < nrow; ++i) {
            if (i == p) continue;
            double wi = w[i];
            if (wi == 0.0) continue;
            double *rowi = Binv + (size_t)i * nrow;
            for (int k = 0; k < nrow; ++k) rowi[k] = rowi[k] - wi * rowp[k];
        }
        pos[basis[p]] = -1;
        basis[p] = q;
        pos[q] = p;
    }
    for (int j = 0; j < ncol; ++j) {
        double v = pos[j] >= 0 ? xB[pos[j]] : 0.0;
        if (j != t_col && v < EPS_ZERO) v = 0.0;
        z[j] = v;
    }
    if (iters_out) *iters_out = it;
    free(Binv); free(xB); free(y); free(w); free(basis); free(pos);
    return rc;
}

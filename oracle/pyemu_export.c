/* TEST INFRASTRUCTURE ONLY: exports pyemu.h helpers for tests/test_pyemu.py. */
#include "pyemu.h"
double fjsp_pysum(const double *x, int n, int mode)
{
    PySum s; pysum_init(&s, mode);
    for (int i = 0; i < n; ++i) pysum_add(&s, x[i]);
    return pysum_result(&s);
}
int fjsp_pyset_order(const int *seq, int n, int *out) { return pyset_order(seq, n, out); }
int fjsp_pyset_intersection_list(const int *a, int na, const int *b, int nb, int *out)
{
    return pyset_intersection_list(a, na, b, nb, out);
}

/* TEST INFRASTRUCTURE ONLY (oracle/): scalar C models of the two CPython behaviours
 * the reference's arithmetic and tie-breaking silently depend on.
 *
 * 1. builtin sum() over floats.  CPython >= 3.12 (the interpreter this image runs
 *    the reference with: 3.12.3) uses Neumaier compensated summation in
 *    Python/bltinmodule.c:builtin_sum_impl; older interpreters add left to right.
 *    Every `sum(...)` in the reference's state_extract / update_parameter /
 *    gap_ave goes through it (e.g. SO_DFJSP.py:89-98,156; class_FJSP.py:159,304).
 *
 * 2. list(set(a) & set(b)) over small ints.  machine_select builds its candidate
 *    lists this way (SO_DFJSP.py:305-306, MO_DFJSP.py:419-427) and max()/min()
 *    return the FIRST extremal element, so ties between machines are broken by the
 *    slot order of CPython's open-addressing set table (Objects/setobject.c:
 *    set_add_entry, set_table_resize, set_intersection).
 *
 * tests/test_pyemu.py checks both against the live interpreter.
 */
#ifndef FJSP_PYEMU_H
#define FJSP_PYEMU_H
#include <math.h>

typedef struct { double f, c; int mode; } PySum;

static inline void pysum_init(PySum *s, int mode) { s->f = 0.0; s->c = 0.0; s->mode = mode; }
static inline void pysum_add(PySum *s, double x)
{
    if (s->mode == 0) { s->f = s->f + x; return; }
    double t = s->f + x;
    if (fabs(s->f) >= fabs(x)) s->c = s->c + ((s->f - t) + x);
    else s->c = s->c + ((x - t) + s->f);
    s->f = t;
}
static inline double pysum_result(const PySum *s)
{
    if (s->mode != 0 && s->c != 0.0 && isfinite(s->c)) return s->f + s->c;
    return s->f;
}

/* Iteration order of set(seq) for distinct ints in [0,32).  Tables of 8 slots hold
 * at most 4 entries (fill*5 < mask*3); the fifth insertion rebuilds into 32 slots,
 * where ints < 32 never collide, so larger sets iterate in ascending order. */
static inline int pyset_order(const int *seq, int n, int *out)
{
    if (n >= 5) {
        unsigned mask = 0;
        for (int i = 0; i < n; ++i) mask |= 1u << seq[i];
        int k = 0;
        for (int v = 0; v < 32; ++v) if (mask >> v & 1) out[k++] = v;
        return k;
    }
    int slot[8];
    for (int i = 0; i < 8; ++i) slot[i] = -1;
    for (int e = 0; e < n; ++e) {
        int v = seq[e];
        unsigned i = (unsigned)v & 7u;
        unsigned perturb = (unsigned)v;
        while (slot[i] >= 0) {          /* LINEAR_PROBES (9) never fit in an 8-slot table */
            perturb >>= 5;
            i = (i * 5u + 1u + perturb) & 7u;
        }
        slot[i] = v;
    }
    int k = 0;
    for (int i = 0; i < 8; ++i) if (slot[i] >= 0) out[k++] = slot[i];
    return k;
}

/* list(set(a) & set(b)): iterate the smaller operand's table (b when sizes tie),
 * keep members of the other, insert into a fresh set, list its table. */
static inline int pyset_intersection_list(const int *a, int na, const int *b, int nb, int *out)
{
    int oa[32], ob[32], keep[32];
    na = pyset_order(a, na, oa);
    nb = pyset_order(b, nb, ob);
    unsigned ma = 0, mb = 0;
    for (int i = 0; i < na; ++i) ma |= 1u << oa[i];
    for (int i = 0; i < nb; ++i) mb |= 1u << ob[i];
    int nk = 0;
    if (nb > na) { for (int i = 0; i < na; ++i) if (mb >> oa[i] & 1) keep[nk++] = oa[i]; }
    else         { for (int i = 0; i < nb; ++i) if (ma >> ob[i] & 1) keep[nk++] = ob[i]; }
    return pyset_order(keep, nk, out);
}
#endif

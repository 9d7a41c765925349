/* TEST INFRASTRUCTURE ONLY (oracle/): a scalar C restatement of the reference's
 * per-instance FJSP environments.  Nothing under oracle/ is imported, linked or
 * executed by the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * It follows the reference's OBJECT MODEL on purpose (explicit job / task lists,
 * list.remove, enumerate positions), so it can be read side by side with:
 *   environments/SO_DFJSP.py            (variant 0)   reset 54-79, state_extract 81-100,
 *                                         update_parameter 102-169, step 171-268,
 *                                         task_select 270-301, machine_select 303-325
 *   environments/MO_DFJSP.py            (variant 1)   reset 58-89, state_extract 91-118,
 *                                         update_parameter 120-187, step 189-298,
 *                                         task_select 300-352, machine_select 354-398,
 *                                         compute_reward 400-417
 *   environments/MO_DFJSP_breakdown.py  (variant 2)   step 189-328 (breakdown scan 204-231)
 *   environments/SO_FJSSP.py            (variant 3)   as variant 0 on class_FJSSP.py
 *   environments/class_FJSP.py / class_MODFJSP.py / class_FJSSP.py
 *                                         reset_parameter, reset_object_add,
 *                                         fluid_model, update_fluid_parameter
 * The CUDA path (csrc/) uses a different, compressed state (per-order counters,
 * bit masks, cached rule choices); agreement between the two is the parity test.
 *
 * Pinned against the reference itself: oracle/make_golden.py steps the unmodified
 * reference classes (through oracle/refshim) and tests/test_oracle_golden.py
 * requires this file to reproduce those trajectories bit for bit.
 *
 * random.choice: the reference draws from Python's global Mersenne Twister.  A
 * vector environment needs one stream per instance, so the random rules take the
 * draw as an explicit 32-bit word per step and pick  seq[word % len(seq)];  the
 * golden generator substitutes the same chooser for random.choice.
 *
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "pyemu.h"

int fjsp_lp_solve_sparse(int ncol, int nrow, const int *colptr, const int *rowidx,
                         const double *vals, const double *b, int t_col,
                         double *z, int *iters_out);

#define FJSP_MAGIC 0x464A5350
#define MAXM 32

enum { V_SO_DFJSP = 0, V_MO_DFJSP = 1, V_MO_BREAKDOWN = 2, V_SO_FJSSP = 3 };

typedef struct { int due, arrive, order, done_tasks; } Job;

typedef struct {
    int r, j;
    int *now; int now_head, now_tail;     /* job_now_list (FIFO of job numbers) */
    int *unp; int unp_len;                /* task_unprocessed_list == job_unprocessed_list */
    int processed;                        /* len(task_processed_list) */
    int fluid_number;                     /* len(job_now_list) at the last arrival */
    double fluid_unp;                     /* fluid_unprocessed_number */
    int fluid_start;                      /* fluid_unprocessed_number_start */
    double rate_sum, time_sum;            /* fluid_rate_sum, fluid_time_sum */
    int nfl; int flm[MAXM];               /* fluid_machine_list (x.items() order) */
    /* dictionaries refreshed by update_parameter for available (r,j) */
    double urgency, delay_e; long long delay_a; int due_min;
    int in_delay_a, in_delay_e;
} Kt;

typedef struct {
    int state; int time_end; int job_r, job_n; int ntasks; int last_task_end;
    long long work;                       /* sum of time_cost, for utilize_rate */
} Machine;

typedef struct {
    int variant, sum_mode;
    int M, K, KT, S, NP, NBD;
    double ddt;
    const int *ntask, *rj_kind, *rj_stage, *nelig, *mt_order, *ptime, *power, *idle_power;
    const int *arrive, *due, *count, *bd_ptr, *bd_start, *bd_end, *pair_order;
    int *blob;
    int *first_rj;                        /* [K] */
    int *narr;                            /* [K] jobs arrived (len(job_arrive_list)) */
    int *ncap;                            /* [K] total jobs over all orders */
    Job **jobs;                           /* [K][ncap] */
    Kt *kt;
    Machine *mach;
    double *unp_mrj, *fl_unp_mrj, *fl_arr_mrj, *frate_mrj; /* [M*KT] indexed m*KT+rj */
    int next_order;
    int step_count, step_time, order_arrive_time, done;
    long long delay_processed, delay_unprocessed, delay_sum, delay_sum_last;
    long long completion, completion_last, energy, energy_last;
    double obs[16], last_obs[16];
    int nobs;
    int lp_solves, lp_iters, error;
} Env;

static int blob_sections(Env *e, const int *b)
{
    if (b[0] != FJSP_MAGIC) return -1;
    e->M = b[2]; e->K = b[3]; e->KT = b[4]; e->S = b[5]; e->NP = b[6]; e->NBD = b[7];
    uint64_t bits = (uint32_t)b[8] | ((uint64_t)(uint32_t)b[9] << 32);
    memcpy(&e->ddt, &bits, 8);
    const int *p = b + 16;
    int M = e->M, K = e->K, KT = e->KT, S = e->S;
    e->ntask = p; p += K;
    e->rj_kind = p; p += KT;
    e->rj_stage = p; p += KT;
    e->nelig = p; p += KT;
    e->mt_order = p; p += KT * M;
    e->ptime = p; p += KT * M;
    e->power = p; p += KT * M;
    e->idle_power = p; p += M;
    e->arrive = p; p += S;
    e->due = p; p += S;
    e->count = p; p += S * K;
    e->bd_ptr = p; p += M + 1;
    e->bd_start = p; p += e->NBD;
    e->bd_end = p; p += e->NBD;
    e->pair_order = p; p += e->NP;
    if (p - b != b[1]) return -2;
    return 0;
}

void *fjsp_oracle_create(const int *blob, int variant, int sum_mode)
{
    Env *e = (Env *)calloc(1, sizeof(Env));
    e->blob = (int *)malloc(sizeof(int) * blob[1]);
    memcpy(e->blob, blob, sizeof(int) * blob[1]);
    if (blob_sections(e, e->blob) != 0 || e->M > MAXM) { free(e->blob); free(e); return NULL; }
    e->variant = variant; e->sum_mode = sum_mode;
    int M = e->M, K = e->K, KT = e->KT, S = e->S;
    e->first_rj = (int *)calloc(K, sizeof(int));
    e->narr = (int *)calloc(K, sizeof(int));
    e->ncap = (int *)calloc(K, sizeof(int));
    e->jobs = (Job **)calloc(K, sizeof(Job *));
    int acc = 0;
    for (int r = 0; r < K; ++r) {
        e->first_rj[r] = acc; acc += e->ntask[r];
        for (int s = 0; s < S; ++s) e->ncap[r] += e->count[s * K + r];
        e->jobs[r] = (Job *)calloc(e->ncap[r] + 1, sizeof(Job));
    }
    e->kt = (Kt *)calloc(KT, sizeof(Kt));
    for (int q = 0; q < KT; ++q) {
        Kt *k = &e->kt[q];
        k->r = e->rj_kind[q]; k->j = e->rj_stage[q];
        k->now = (int *)calloc(e->ncap[k->r] + 1, sizeof(int));
        k->unp = (int *)calloc(e->ncap[k->r] + 1, sizeof(int));
    }
    e->mach = (Machine *)calloc(M, sizeof(Machine));
    e->unp_mrj = (double *)calloc((size_t)M * KT, sizeof(double));
    e->fl_unp_mrj = (double *)calloc((size_t)M * KT, sizeof(double));
    e->fl_arr_mrj = (double *)calloc((size_t)M * KT, sizeof(double));
    e->frate_mrj = (double *)calloc((size_t)M * KT, sizeof(double));
    e->nobs = (variant == V_MO_DFJSP || variant == V_MO_BREAKDOWN) ? 15 : 10;
    return e;
}

void fjsp_oracle_destroy(void *h)
{
    Env *e = (Env *)h;
    if (!e) return;
    for (int r = 0; r < e->K; ++r) free(e->jobs[r]);
    for (int q = 0; q < e->KT; ++q) { free(e->kt[q].now); free(e->kt[q].unp); }
    free(e->jobs); free(e->kt); free(e->mach); free(e->first_rj); free(e->narr); free(e->ncap);
    free(e->unp_mrj); free(e->fl_unp_mrj); free(e->fl_arr_mrj); free(e->frate_mrj);
    free(e->blob); free(e);
}

static inline int is_mo(const Env *e) { return e->variant == V_MO_DFJSP || e->variant == V_MO_BREAKDOWN; }
static inline int elig(const Env *e, int q, int m) { return e->ptime[q * e->M + m] > 0; }
static inline int now_len(const Kt *k) { return k->now_tail - k->now_head; }

/* ------------------------------------------------------------ fluid model ---- */
/* class_FJSP.py:256-290 fluid_model + 292-316 update_fluid_parameter.
 * Canonical LP: columns = pairs ((r,j) ascending, m ascending) then t;
 * rows = capacity m=0..M-1, demand (r,j) ascending, precedence (r,j) ascending. */
static int solve_fluid(Env *e)
{
    int M = e->M, KT = e->KT;
    int *col_of = (int *)malloc(sizeof(int) * KT * M);
    int np = 0;
    for (int q = 0; q < KT; ++q)
        for (int m = 0; m < M; ++m) col_of[q * M + m] = elig(e, q, m) ? np++ : -1;
    int *prec_row = (int *)malloc(sizeof(int) * KT);
    int nprec = 0;
    for (int q = 0; q < KT; ++q) {
        prec_row[q] = -1;
        int r = e->kt[q].r, j = e->kt[q].j;
        if (j + 1 < e->ntask[r] && e->kt[q + 1].fluid_number == 0) prec_row[q] = M + KT + nprec++;
    }
    int ncol = np + 1, nrow = M + KT + nprec, t_col = np;
    int *colptr = (int *)malloc(sizeof(int) * (ncol + 1));
    int *rowidx = (int *)malloc(sizeof(int) * (4 * np + KT));
    double *vals = (double *)malloc(sizeof(double) * (4 * np + KT));
    double *b = (double *)calloc(nrow, sizeof(double));
    double *z = (double *)calloc(ncol, sizeof(double));
    for (int m = 0; m < M; ++m) b[m] = 1.0;
    int nz = 0, c = 0;
    for (int q = 0; q < KT; ++q) {
        int j = e->kt[q].j;
        for (int m = 0; m < M; ++m) {
            if (!elig(e, q, m)) continue;
            double rate = 1.0 / (double)e->ptime[q * M + m];
            colptr[c++] = nz;
            rowidx[nz] = m; vals[nz++] = 1.0;
            rowidx[nz] = M + q; vals[nz++] = -(rate / (double)e->kt[q].fluid_start);
            if (j > 0 && prec_row[q - 1] >= 0) { rowidx[nz] = prec_row[q - 1]; vals[nz++] = rate; }
            if (prec_row[q] >= 0) { rowidx[nz] = prec_row[q]; vals[nz++] = -rate; }
        }
    }
    colptr[c++] = nz;
    for (int q = 0; q < KT; ++q) { rowidx[nz] = M + q; vals[nz++] = 1.0; }
    colptr[c] = nz;
    int iters = 0;
    int rc = fjsp_lp_solve_sparse(ncol, nrow, colptr, rowidx, vals, b, t_col, z, &iters);
    e->lp_solves++; e->lp_iters += iters;
    if (rc != 0) e->error |= 1;
    /* reset_fluid_parameter + update_fluid_parameter, x.items() order = pair_order */
    for (int q = 0; q < KT; ++q) e->kt[q].nfl = 0;
    for (size_t i = 0; i < (size_t)M * KT; ++i) e->frate_mrj[i] = 0.0;
    PySum *rs = (PySum *)malloc(sizeof(PySum) * KT);
    for (int q = 0; q < KT; ++q) pysum_init(&rs[q], e->sum_mode);
    for (int i = 0; i < e->NP; ++i) {
        int q = e->pair_order[i] / M, m = e->pair_order[i] % M;
        double x = z[col_of[q * M + m]];
        double fr = x * (1.0 / (double)e->ptime[q * M + m]);
        e->frate_mrj[m * KT + q] = fr;
        pysum_add(&rs[q], fr);
        if (x != 0.0) e->kt[q].flm[e->kt[q].nfl++] = m;
    }
    for (int q = 0; q < KT; ++q) {
        e->kt[q].rate_sum = pysum_result(&rs[q]);
        e->kt[q].time_sum = 1.0 / e->kt[q].rate_sum;
    }
    for (int m = 0; m < M; ++m)
        for (int q = 0; q < KT; ++q) {
            if (!elig(e, q, m)) continue;
            double arr = (double)e->kt[q].fluid_start * e->frate_mrj[m * KT + q] / e->kt[q].rate_sum;
            e->fl_arr_mrj[m * KT + q] = arr;
            e->unp_mrj[m * KT + q] = arr;
            e->fl_unp_mrj[m * KT + q] = arr;
        }
    free(rs); free(col_of); free(prec_row); free(colptr); free(rowidx); free(vals); free(b); free(z);
    return rc;
}

/* class_FJSP.py:218-254 reset_object_add */
static void order_arrives(Env *e, int s)
{
    int K = e->K;
    for (int r = 0; r < K; ++r) {
        int n0 = e->narr[r], cnt = e->count[s * K + r];
        int r_due = 0;
        if (e->variant == V_SO_FJSSP) /* class_FJSSP.py:214-215 */
            r_due = (int)rint((double)((long long)e->due[s] * e->ntask[r]) / (double)cnt);
        for (int n = n0; n < n0 + cnt; ++n) {
            Job *jb = &e->jobs[r][n];
            jb->due = e->due[s];
            if (e->variant == V_SO_FJSSP) /* class_FJSSP.py:218 */
                jb->due = (int)rint((double)((long long)r_due * n) / (double)cnt);
            jb->arrive = e->arrive[s]; jb->order = s; jb->done_tasks = 0;
            Kt *k0 = &e->kt[e->first_rj[r]];
            k0->now[k0->now_tail++] = n;
            for (int j = 0; j < e->ntask[r]; ++j) {
                Kt *k = &e->kt[e->first_rj[r] + j];
                k->unp[k->unp_len++] = n;
            }
        }
        e->narr[r] = n0 + cnt;
    }
    for (int q = 0; q < e->KT; ++q) {
        Kt *k = &e->kt[q];
        k->fluid_number = now_len(k);
        k->fluid_unp = (double)k->unp_len;
        k->fluid_start = k->unp_len;
    }
    solve_fluid(e);
}

/* ------------------------------------------------------------ availability ---- */
static unsigned idle_mask(const Env *e)
{
    unsigned mk = 0;
    for (int m = 0; m < e->M; ++m) if (e->mach[m].state == 0) mk |= 1u << m;
    return mk;
}
static int avail(const Env *e, int q, unsigned idle)
{
    if (now_len(&e->kt[q]) <= 0) return 0;
    for (int m = 0; m < e->M; ++m) if ((idle >> m & 1) && elig(e, q, m)) return 1;
    return 0;
}
static int fluid_avail(const Env *e, int q, unsigned idle)
{
    if (now_len(&e->kt[q]) <= 0) return 0;
    for (int i = 0; i < e->kt[q].nfl; ++i) if (idle >> e->kt[q].flm[i] & 1) return 1;
    return 0;
}
static int count_avail(const Env *e, int fluid)
{
    unsigned idle = idle_mask(e);
    int c = 0;
    for (int q = 0; q < e->KT; ++q) c += fluid ? fluid_avail(e, q, idle) : avail(e, q, idle);
    return c;
}

/* ------------------------------------------------------------ update_parameter -- */
static void update_parameter(Env *e, double *out4)
{
    long long da = 0, de = 0, tn = 0, ja = 0, je = 0, jn = 0;
    e->delay_unprocessed = 0;
    int t = e->step_time;
    double td = (double)t;
    unsigned idle = idle_mask(e);
    for (int r = 0; r < e->K; ++r) {
        Kt *kl = &e->kt[e->first_rj[r] + e->ntask[r] - 1];
        jn += kl->unp_len;
        for (int i = 0; i < kl->unp_len; ++i) {
            Job *jb = &e->jobs[r][kl->unp[i]];
            if (t > jb->due) { ja++; e->delay_unprocessed += t - jb->due; }
            if (td + kl->time_sum * (double)(i + 1) > (double)jb->due) je++;
        }
    }
    for (int q = 0; q < e->KT; ++q) {
        Kt *k = &e->kt[q];
        int residue = k->unp_len;
        tn += residue;
        int a = 0, ec = 0;
        long long max_a = 0; double max_e = 0.0;
        PySum se; pysum_init(&se, e->sum_mode);
        for (int i = 0; i < residue; ++i) {
            Job *jb = &e->jobs[k->r][k->unp[i]];
            double est = td + k->time_sum * (double)(i + 1);
            if (t > jb->due) a++;
            if (est > (double)jb->due) ec++;
            long long va = (long long)t - jb->due;
            double ve = est - (double)jb->due;
            if (i == 0 || va > max_a) max_a = va;
            if (i == 0 || ve > max_e) max_e = ve;
            pysum_add(&se, ve);
        }
        da += a; de += ec;
        k->in_delay_a = k->in_delay_e = 0;
        if (avail(e, q, idle)) {
            if (a > 0) { k->in_delay_a = 1; k->delay_a = max_a; }
            if (ec > 0) { k->in_delay_e = 1; k->delay_e = max_e; }
            k->urgency = pysum_result(&se) / (double)residue;
            int dm = 0;
            for (int i = k->now_head; i < k->now_tail; ++i) {
                int d = e->jobs[k->r][k->now[i]].due;
                if (i == k->now_head || d < dm) dm = d;
            }
            k->due_min = dm;
        }
    }
    if (!e->done) {
        out4[0] = (double)da / (double)tn; out4[1] = (double)de / (double)tn;
        out4[2] = (double)ja / (double)jn; out4[3] = (double)je / (double)jn;
    } else {
        out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
    }
}

static double kt_gap(const Kt *k) { return (double)k->unp_len - k->fluid_unp; }
static double mach_gap_rj(const Env *e, int m, int q)
{
    return e->unp_mrj[m * e->KT + q] - e->fl_unp_mrj[m * e->KT + q];
}
static double mach_gap_ave(const Env *e, int m)
{
    PySum s; pysum_init(&s, e->sum_mode);
    int n = 0;
    for (int q = 0; q < e->KT; ++q) if (elig(e, q, m)) { pysum_add(&s, mach_gap_rj(e, m, q)); n++; }
    return pysum_result(&s) / (double)n;
}

/* ------------------------------------------------------------ state_extract ---- */
static void state_extract(Env *e, double *obs)
{
    int M = e->M, KT = e->KT;
    long long tsum = 0;
    for (int m = 0; m < M; ++m) tsum += e->mach[m].time_end;
    double ct_ave = (double)tsum / (double)M;
    PySum s; pysum_init(&s, e->sum_mode);
    for (int m = 0; m < M; ++m) pysum_add(&s, pow((double)e->mach[m].time_end - ct_ave, 2.0));
    double ct_std = sqrt(pysum_result(&s) / (double)M);
    pysum_init(&s, e->sum_mode);
    for (int q = 0; q < KT; ++q)
        pysum_add(&s, (double)e->kt[q].processed / (double)(e->kt[q].unp_len + e->kt[q].processed));
    double cro_ave = pysum_result(&s) / (double)KT;
    pysum_init(&s, e->sum_mode);
    for (int q = 0; q < KT; ++q)
        pysum_add(&s, pow((double)e->kt[q].processed / (double)(e->kt[q].unp_len + e->kt[q].processed) - cro_ave, 2.0));
    double cro_std = sqrt(pysum_result(&s) / (double)KT);
    pysum_init(&s, e->sum_mode);
    for (int q = 0; q < KT; ++q) pysum_add(&s, kt_gap(&e->kt[q]) / (double)e->kt[q].fluid_start);
    double gap_ave = pysum_result(&s) / (double)KT;
    pysum_init(&s, e->sum_mode);
    for (int q = 0; q < KT; ++q)
        pysum_add(&s, pow(kt_gap(&e->kt[q]) / (double)e->kt[q].fluid_start - gap_ave, 2.0));
    double gap_std = sqrt(pysum_result(&s) / (double)KT);
    double d4[4];
    if (is_mo(e)) {
        double ratio_idle = (double)count_avail(e, 1) / ((double)count_avail(e, 0) + 1e-08);
        double gm[MAXM];
        pysum_init(&s, e->sum_mode);
        for (int m = 0; m < M; ++m) { gm[m] = mach_gap_ave(e, m); pysum_add(&s, gm[m]); }
        double gm_ave = pysum_result(&s) / (double)M;
        pysum_init(&s, e->sum_mode);
        for (int m = 0; m < M; ++m) pysum_add(&s, pow(gm[m] - gm_ave, 2.0));
        double gm_std = sqrt(pysum_result(&s) / (double)M);
        update_parameter(e, d4);
        obs[0] = e->ddt; obs[1] = (double)M; obs[2] = (double)e->S; obs[3] = ct_std; obs[4] = ratio_idle;
        obs[5] = cro_ave; obs[6] = cro_std; obs[7] = gap_ave; obs[8] = gap_std;
        obs[9] = gm_ave; obs[10] = gm_std;
        obs[11] = d4[0]; obs[12] = d4[1]; obs[13] = d4[2]; obs[14] = d4[3];
    } else {
        update_parameter(e, d4);
        obs[0] = (double)M; obs[1] = ct_std; obs[2] = cro_ave; obs[3] = cro_std;
        obs[4] = gap_ave; obs[5] = gap_std;
        obs[6] = d4[0]; obs[7] = d4[1]; obs[8] = d4[2]; obs[9] = d4[3];
    }
}

static void emit_state(Env *e, double *state_out)
{
    for (int i = 0; i < e->nobs; ++i) {
        state_out[i] = e->obs[i];
        state_out[e->nobs + i] = e->obs[i] - e->last_obs[i];
    }
}

/* ------------------------------------------------------------ reset ---------- */
/* SO_DFJSP.py:54-79 / MO_DFJSP.py:58-89 on top of reset_parameter (class_FJSP.py:186-203).
 * Three reference quirks are kept so that a re-reset of a used environment (what the
 * agents do every episode, and what the fused auto-reset does) matches:
 *   - machine.state is NOT cleared (reset_parameter assigns `machine_state`);
 *   - order_arrive_time keeps its value from the previous episode;
 *   - self.done is cleared only after the two state_extract() calls, so after a
 *     finished episode the four delay rates of the reset observation are 0. */
int fjsp_oracle_reset(void *h, double *state_out)
{
    Env *e = (Env *)h;
    for (int r = 0; r < e->K; ++r) e->narr[r] = 0;
    for (int q = 0; q < e->KT; ++q) {
        Kt *k = &e->kt[q];
        k->now_head = k->now_tail = 0; k->unp_len = 0; k->processed = 0;
        k->in_delay_a = k->in_delay_e = 0;
    }
    for (int m = 0; m < e->M; ++m) {
        Machine *mc = &e->mach[m];
        mc->time_end = 0; mc->job_r = mc->job_n = -1; mc->ntasks = 0; mc->last_task_end = 0; mc->work = 0;
    }
    e->next_order = 0;
    order_arrives(e, e->next_order++);
    e->delay_sum_last = e->delay_sum = e->delay_processed = e->delay_unprocessed = 0;
    e->completion = e->completion_last = e->energy = e->energy_last = 0;
    e->step_count = 0; e->step_time = 0;
    state_extract(e, e->last_obs);
    state_extract(e, e->obs);
    emit_state(e, state_out);
    e->done = 0;
    return e->error;
}

/* ------------------------------------------------------------ rule helpers ---- */
static int selectable(const Env *e, int q, int fluid, int *out)
{
    int idle[MAXM], ni = 0;
    for (int m = 0; m < e->M; ++m) if (e->mach[m].state == 0) idle[ni++] = m;
    if (fluid) return pyset_intersection_list(idle, ni, e->kt[q].flm, e->kt[q].nfl, out);
    return pyset_intersection_list(idle, ni, e->mt_order + q * e->M, e->nelig[q], out);
}
static long long energy_mrj(const Env *e, int m, int q)
{
    return (long long)e->power[q * e->M + m] * e->ptime[q * e->M + m];
}
/* MO_DFJSP.py:429-451 time_min_rj / time_min_fluid_rj / energy_min_rj / energy_min_fluid_rj */
static long long min_over_selectable(const Env *e, int q, int fluid, int energy)
{
    int lst[MAXM];
    int n = selectable(e, q, fluid, lst);
    long long best = 0;
    for (int i = 0; i < n; ++i) {
        long long v = energy ? energy_mrj(e, lst[i], q) : e->ptime[q * e->M + lst[i]];
        if (i == 0 || v < best) best = v;
    }
    return best;
}

#define ARGMAX_D(list, n, expr)  do { int bi_ = -1; double bv_ = 0; for (int i_ = 0; i_ < (n); ++i_) { int x = (list)[i_]; double v_ = (expr); if (bi_ < 0 || v_ > bv_) { bi_ = x; bv_ = v_; } } sel = bi_; } while (0)
#define ARGMIN_D(list, n, expr)  do { int bi_ = -1; double bv_ = 0; for (int i_ = 0; i_ < (n); ++i_) { int x = (list)[i_]; double v_ = (expr); if (bi_ < 0 || v_ < bv_) { bi_ = x; bv_ = v_; } } sel = bi_; } while (0)

static int task_select(Env *e, int rule, uint32_t rnd)
{
    int KT = e->KT;
    unsigned idle = idle_mask(e);
    int *av = (int *)malloc(sizeof(int) * KT * 4);
    int *fav = av + KT, *dl_e = av + 2 * KT, *dl_a = av + 3 * KT;
    int nav = 0, nfav = 0, ne = 0, na = 0;
    for (int q = 0; q < KT; ++q) {
        if (avail(e, q, idle)) av[nav++] = q;
        if (fluid_avail(e, q, idle)) fav[nfav++] = q;
        if (e->kt[q].in_delay_e) dl_e[ne++] = q;
        if (e->kt[q].in_delay_a) dl_a[na++] = q;
    }
    int sel = -1;
    int mo = is_mo(e);
    int *fl_or_av = nfav ? fav : av; int nfl_or_av = nfav ? nfav : nav;
    if (nav == 0) { free(av); return -1; }
    switch (rule) {
    case 1: if (ne == 0) ARGMAX_D(av, nav, e->kt[x].urgency); else ARGMAX_D(dl_e, ne, e->kt[x].delay_e); break;
    case 2: if (na == 0) ARGMAX_D(av, nav, e->kt[x].urgency); else ARGMAX_D(dl_a, na, (double)e->kt[x].delay_a); break;
    case 3: ARGMAX_D(fl_or_av, nfl_or_av, kt_gap(&e->kt[x])); break;
    case 4: ARGMAX_D(fl_or_av, nfl_or_av, e->kt[x].urgency); break;
    case 5: ARGMIN_D(fl_or_av, nfl_or_av, (double)e->kt[x].due_min); break;
    default:
        if (!mo) { if (rule == 6) sel = av[rnd % (uint32_t)nav]; break; }
        switch (rule) {
        case 6: ARGMIN_D(av, nav, (double)e->kt[x].due_min); break;
        case 7: if (nfav == 0) ARGMIN_D(av, nav, (double)min_over_selectable(e, x, 0, 1));
                else ARGMIN_D(fav, nfav, (double)min_over_selectable(e, x, 1, 1)); break;
        case 8: ARGMIN_D(av, nav, (double)min_over_selectable(e, x, 0, 1)); break;
        case 9: if (nfav == 0) ARGMIN_D(av, nav, (double)min_over_selectable(e, x, 0, 0));
                else ARGMIN_D(fav, nfav, (double)min_over_selectable(e, x, 1, 0)); break;
        case 10: ARGMIN_D(av, nav, (double)min_over_selectable(e, x, 0, 0)); break;
        case 11: sel = fl_or_av[rnd % (uint32_t)nfl_or_av]; break;
        case 12: sel = av[rnd % (uint32_t)nav]; break;
        }
    }
    free(av);
    return sel;
}

static int machine_select(Env *e, int rule, int q, uint32_t rnd)
{
    int sl[MAXM], fl[MAXM];
    int ns = selectable(e, q, 0, sl), nf = selectable(e, q, 1, fl);
    int *fs = nf ? fl : sl; int nfs = nf ? nf : ns;
    int sel = -1, M = e->M;
    if (ns == 0) return -1;
    if (!is_mo(e)) {
        switch (rule) {
        case 1: ARGMAX_D(fs, nfs, mach_gap_rj(e, x, q)); break;
        case 2: ARGMAX_D(sl, ns, mach_gap_rj(e, x, q)); break;
        case 3: ARGMIN_D(sl, ns, (double)e->ptime[q * M + x]); break;
        case 4: ARGMAX_D(fs, nfs, mach_gap_ave(e, x)); break;
        case 5: sel = sl[rnd % (uint32_t)ns]; break;
        }
        return sel;
    }
    switch (rule) {
    case 1: ARGMAX_D(fs, nfs, mach_gap_rj(e, x, q)); break;
    case 2: ARGMIN_D(fs, nfs, (double)e->ptime[q * M + x]); break;
    case 3: ARGMIN_D(sl, ns, (double)e->ptime[q * M + x]); break;
    case 4: ARGMAX_D(fs, nfs, mach_gap_ave(e, x)); break;
    case 5: ARGMIN_D(fs, nfs, (double)energy_mrj(e, x, q)); break;
    case 6: ARGMIN_D(sl, ns, (double)energy_mrj(e, x, q)); break;
    case 7: ARGMIN_D(fs, nfs, (double)e->idle_power[x]); break;
    case 8: ARGMIN_D(sl, ns, (double)e->idle_power[x]); break;
    case 9: sel = fs[rnd % (uint32_t)nfs]; break;
    case 10: sel = sl[rnd % (uint32_t)ns]; break;
    }
    return sel;
}

/* ------------------------------------------------------------ step ----------- */
/* rec[8] = {rj, kind, stage, job_number, machine, time_begin, time_end, machine_time_end} */
int fjsp_oracle_step(void *h, int task_rule, int machine_rule, uint32_t rnd_task, uint32_t rnd_machine,
                     int reward_policy, double completion, double tardiness, double energy_norm,
                     double *state_out, double *reward_out, int *done_out, int *rec)
{
    Env *e = (Env *)h;
    int M = e->M, KT = e->KT;
    int q = task_select(e, task_rule + 1, rnd_task);
    if (q < 0) { e->error |= 2; return e->error; }
    int m = machine_select(e, machine_rule + 1, q, rnd_machine);
    if (m < 0) { e->error |= 4; return e->error; }
    Kt *k = &e->kt[q];
    int n = k->now[k->now_head];
    Job *jb = &e->jobs[k->r][n];
    Machine *mc = &e->mach[m];
    int dur = e->ptime[q * M + m];
    int t_begin = e->step_time, t_end = e->step_time + dur, m_end = t_end;
    if (e->variant == V_MO_BREAKDOWN) { /* MO_DFJSP_breakdown.py:204-231 */
        int cur = e->step_time;
        for (int i = e->bd_ptr[m]; i < e->bd_ptr[m + 1]; ++i) {
            int bs = e->bd_start[i], be = e->bd_end[i];
            if (bs <= cur && cur < be) { int d = be - cur; t_begin += d; t_end += d; m_end = t_end; }
            else if (cur < bs && bs < t_end) { t_end += be - bs; m_end = t_end; }
            else if (bs == t_end) { m_end += be - bs; }
            else if (bs > t_end) break;
        }
    }
    /* job / kind-task bookkeeping */
    jb->done_tasks++;
    k->now_head++;
    for (int i = 0; i < k->unp_len; ++i)
        if (k->unp[i] == n) { memmove(k->unp + i, k->unp + i + 1, sizeof(int) * (k->unp_len - i - 1)); break; }
    k->unp_len--; k->processed++;
    /* machine */
    int prev_task_end = mc->last_task_end, had = mc->ntasks;
    mc->state = 1; mc->time_end = m_end; mc->ntasks++; mc->last_task_end = t_end;
    mc->job_r = k->r; mc->job_n = n; mc->work += t_end - t_begin;
    e->unp_mrj[m * KT + q] = e->unp_mrj[m * KT + q] - 1.0;
    if (is_mo(e)) {
        if (t_end > e->completion) e->completion = t_end;
        e->energy += energy_mrj(e, m, q);
        if (had >= 1) e->energy += (long long)(e->step_time - prev_task_end) * e->idle_power[m];
    }
    if (jb->done_tasks == e->ntask[k->r]) {
        long long late = (long long)t_end - jb->due;
        e->delay_processed += late > 0 ? late : 0;
    }
    if (rec) { rec[0] = q; rec[1] = k->r; rec[2] = k->j; rec[3] = n; rec[4] = m; rec[5] = t_begin; rec[6] = t_end; rec[7] = m_end; }
    /* advance the clock while nothing can be dispatched */
    while (count_avail(e, 0) == 0) {
        int tmin = 0, found = 0;
        for (int i = 0; i < M; ++i)
            if (e->mach[i].time_end > e->step_time && (!found || e->mach[i].time_end < tmin)) { tmin = e->mach[i].time_end; found = 1; }
        if (!found) { e->error |= 8; break; }
        e->step_time = tmin;
        for (int i = 0; i < M; ++i) {
            Machine *mi = &e->mach[i];
            if (mi->time_end == e->step_time && mi->job_r >= 0) {
                Job *j2 = &e->jobs[mi->job_r][mi->job_n];
                if (j2->done_tasks < e->ntask[mi->job_r]) {
                    Kt *k2 = &e->kt[e->first_rj[mi->job_r] + j2->done_tasks];
                    k2->now[k2->now_tail++] = mi->job_n;
                }
            }
        }
        long long left = 0;
        for (int r = 0; r < e->K; ++r) left += e->kt[e->first_rj[r] + e->ntask[r] - 1].unp_len;
        if (e->next_order < e->S && e->arrive[e->next_order] <= e->step_time) {
            int s = e->next_order++;
            order_arrives(e, s);
            e->order_arrive_time = e->arrive[s];
        } else if (e->next_order < e->S && left == 0) {
            int s = e->next_order++;
            order_arrives(e, s);
            e->order_arrive_time = e->arrive[s];
            e->step_time = e->order_arrive_time;
        }
        for (int i = 0; i < M; ++i) if (e->mach[i].time_end <= e->step_time) e->mach[i].state = 0;
        double gap_time = (double)(e->step_time - e->order_arrive_time);
        for (int x = 0; x < KT; ++x)
            e->kt[x].fluid_unp = (double)e->kt[x].fluid_start - e->kt[x].rate_sum * gap_time;
        for (int i = 0; i < M; ++i)
            for (int x = 0; x < KT; ++x)
                if (elig(e, x, i))
                    e->fl_unp_mrj[i * KT + x] = e->fl_arr_mrj[i * KT + x] - gap_time * e->frate_mrj[i * KT + x];
        left = 0;
        for (int r = 0; r < e->K; ++r) left += e->kt[e->first_rj[r] + e->ntask[r] - 1].unp_len;
        if (e->next_order >= e->S && left == 0) { e->done = 1; break; }
    }
    e->step_count++;
    memcpy(e->last_obs, e->obs, sizeof(e->obs));
    state_extract(e, e->obs);
    emit_state(e, state_out);
    e->delay_sum = e->delay_processed + e->delay_unprocessed;
    double rew = 0.0;
    if (!is_mo(e)) rew = -(double)(e->delay_sum - e->delay_sum_last);
    else if (reward_policy == 0) rew = (double)(e->completion_last - e->completion);
    else if (reward_policy == 1) rew = (double)(e->delay_sum_last - e->delay_sum);
    else if (reward_policy == 2) rew = (double)(e->energy_last - e->energy);
    else if (reward_policy == 3) {
        double a = (double)(e->completion_last - e->completion) / completion;
        double c = (double)(e->energy_last - e->energy) / energy_norm;
        if (tardiness > 0) rew = a + (double)(e->delay_sum_last - e->delay_sum) / tardiness + c;
        else rew = a + c;
    }
    e->delay_sum_last = e->delay_sum; e->completion_last = e->completion; e->energy_last = e->energy;
    *reward_out = rew; *done_out = e->done;
    return e->error;
}

/* info[0..9]: step_time, step_count, completion, delay_sum, energy, lp_solves, lp_iters, error, done, next_order */
void fjsp_oracle_info(void *h, long long *info)
{
    Env *e = (Env *)h;
    info[0] = e->step_time; info[1] = e->step_count; info[2] = e->completion; info[3] = e->delay_sum;
    info[4] = e->energy; info[5] = e->lp_solves; info[6] = e->lp_iters; info[7] = e->error; info[8] = e->done;
    info[9] = e->next_order;
}

int fjsp_oracle_done(void *h) { return ((Env *)h)->done; }

/* machine completion times, for makespan checks in the SO variants */
void fjsp_oracle_machine_end(void *h, int *out)
{
    Env *e = (Env *)h;
    for (int m = 0; m < e->M; ++m) out[m] = e->mach[m].time_end;
}

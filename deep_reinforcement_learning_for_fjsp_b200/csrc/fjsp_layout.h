// Layout of the device-resident vector environment (shared by host and device code).
//
// Two tables live in HBM:
//   * the INSTANCE table: one read-only int32 record per distinct problem instance
//     (shared by every environment copy that plays it; stays hot in L2);
//   * the ENV table: one mutable record per environment copy, contiguous per env so
//     that the warp that owns an env streams it with coalesced loads.
// All offsets are computed once on the host (fjsp_host.h) from the batch maxima and
// passed to the kernels by value.
#pragma once
#include <stdint.h>

#define FJSP_MAGIC 0x464A5350
#define FJSP_MAX_M 32
#define FJSP_MAX_KT 256
#define FJSP_MAX_S 16

enum { FJSP_SO_DFJSP = 0, FJSP_MO_DFJSP = 1, FJSP_MO_BREAKDOWN = 2, FJSP_SO_FJSSP = 3 };

struct FjDims {
    int Mx, Kx, KTx, Sx, NBDx, NJx;   // batch maxima
    int NFx;                          // fluid-pair slots per env (max LP rows)
    int KTW;                          // 32-bit words per operation-type mask
    int Rx, NPx;                      // LP: max rows, max pair columns
    int NWx;                          // SO_FJSSP: 32-bit words per unprocessed-job set (0 otherwise)
};

// instance record: word (int32) offsets
struct FjInstOff {
    int hdr;       // [12]: M, K, KT, S, NJ, ddt_lo, ddt_hi, NP, operations per episode, 0, 0, 0
    int ntask;     // [Kx]
    int first;     // [Kx]  first operation type of a kind
    int jobbase;   // [Kx]  offset of a kind's jobs in the per-env link array
    int rjkind;    // [KTx]
    int rjstage;   // [KTx]
    int rjlast;    // [KTx] 1 if last stage of its kind
    int elig;      // [KTx] machine bit mask
    int nelig;     // [KTx]
    int mtset;     // [KTx*Mx] iteration order of set(machine_tuple) (CPython slot order)
    int mtpack;    // [KTx] the same order packed one byte per member when the type has at most 4 machines (else 0)
    int poord;     // [KTx*Mx] machines of the type in x.items() (pair) order
    int ptime;     // [KTx*Mx]
    int energy;    // [KTx*Mx] power * time
    int idlep;     // [Mx]
    int arrive;    // [Sx]
    int due;       // [Sx]
    int count;     // [Sx*Kx]
    int cum;       // [(Sx+1)*Kx] jobs of kind r arrived before order s
    int bdptr;     // [Mx+1]
    int bds;       // [NBDx]
    int bde;       // [NBDx]
    int mnkt;      // [Mx] operation types a machine can process (len(kind_task_tuple))
    int colbase;   // [KTx] first LP column of an operation type (prefix of popcount(elig))
    int colqm;     // [NPx] operation type << 8 | machine of every LP column (canonical order: type ascending, machine ascending)
    int colrate;   // [2*NPx] double 1.0 / processing time of every LP column (even word offset: 8-byte aligned)
    int hotw;      // words of the record's hot head (everything a step reads except the per-pair tables)
    int stride;    // words per instance
};

// env record: BYTE offsets
struct FjEnvOff {
    int scal;      // int32[FJ_S_COUNT] scalars, see FJ_S_* below
    int obs;       // double[16] v(t)
    int obs2;      // double[16] staging for v(t+1)
    int gapave;    // double[Mx] cached machine gap_ave
    int urg;       // double[KTx] delivery urgency of an available operation type (rule key)
    int maxe;      // double[KTx] its largest estimated delay (rule key)
    int avmask;    // uint32[KTW] available operation types
    int favmask;   // uint32[KTW] fluid-available operation types
    int demask;    // uint32[KTW] available with an estimated-late operation
    int damask;    // uint32[KTW] available with an actually-late operation
    int mend;      // int32[Mx] machine completion time
    int mlast;     // int32[Mx] end time of the machine's previous operation
    int h_elig;    // uint32[KTx] copy of the instance's eligible-machine masks
    int h_rjinfo;  // uint16[KTx] kind << 8 | stage << 1 | last-stage flag
    int h_mtpack;  // uint32[KTx] copy of the instance's packed set(machine_tuple) orders
    int h_fo;      // uint32[KTx] iteration order of set(fluid machines of the type), packed (at most 4 members), set at order arrival
    int h_due;     // int32[Sx]   copy of the orders' due dates
    int h_cum;     // int32[(Sx+1)*Kx] copy of the cumulative job counts (job number -> order)
    int h_jobbase; // int32[Kx]
    int mF;        // double[Mx] sum of the machine's fluid rates since the last arrival
    int invnkt;    // double[Mx] 1 / (operation types the machine can process): observation features only
    int mD;        // int32[Mx] dispatches on the machine since the last arrival
    int mjob;      // int32[Mx] (rj << 16 | job number) of the job on the machine, -1 none
    int qhead;     // uint16[KTx] stage>0 waiting queue (linked through `next`)
    int qtail;     // uint16[KTx]
    int qlen;      // uint16[KTx]
    int proc;      // int32[KTx] operations dispatched this episode
    int fstart;    // int32[KTx] unprocessed count at the last order arrival
    int flmask;    // uint32[KTx] machines with non-zero fluid rate
    int rsum;      // double[KTx] fluid_rate_sum
    int tsum;      // double[KTx] fluid_time_sum
    int cntunp;    // uint16[KTx*Sx] unprocessed operations per order
    int cntnow;    // uint16[KTx*Sx] waiting jobs per order
    int pk;        // uint16[KTx*Mx] dispatches of type on machine since the last arrival
    int slot;      // uint16[KTx*Mx] fluid slot of the pair, 0xFFFF = none
    int fu;        // double[NFx] unprocessed_rj_dict of fluid pairs
    int fa;        // double[NFx] fluid_unprocessed_rj_arrival_dict
    int ff;        // double[NFx] fluid_process_rate_rj_dict
    int next;      // uint16[NJx] queue links, indexed jobbase[r] + n
    int unpmask;   // uint32[KTx*NWx] SO_FJSSP: unprocessed jobs of an operation type, one bit per job
    int duejob;    // int32[NJx]      SO_FJSSP: due date of every job
    int mindue;    // int32[KTx]      SO_FJSSP: smallest due date among an operation type's unprocessed jobs
    int hot;       // bytes of the record's hot part (everything but `next`), multiple of 16
    int stride;    // bytes per env (multiple of 16)
};

// scalar slots (int32 index into scal[])
enum {
    FJ_S_TIME = 0, FJ_S_ARRTIME = 1, FJ_S_NEXTORDER = 2, FJ_S_DONE = 3, FJ_S_STEPS = 4, FJ_S_BUSY = 5,
    FJ_S_ERROR = 6, FJ_S_HASTASK = 7, FJ_S_COMPLETION = 8, FJ_S_COMPLETION_LAST = 9, FJ_S_EPISODES = 10,
    FJ_S_LPSOLVES = 11, FJ_S_LPITERS = 12, FJ_S_NFL = 13, FJ_S_PHASE = 14, FJ_S_LPSLOT = 15,
    // 64-bit values occupy two slots (even index)
    FJ_S_ENERGY = 16, FJ_S_ENERGY_LAST = 18, FJ_S_DELAY_PROC = 20, FJ_S_DELAY_LAST = 22, FJ_S_DELAY_UNPROC = 24,
    FJ_S_GAPTIME = 26 /* double */, FJ_S_TT = 28, FJ_S_WASDONE = 29, FJ_S_NAV = 30, FJ_S_NFAV = 31,
    FJ_S_LEFT = 32 /* unprocessed last-stage operations = jobs not fully dispatched */,
    FJ_S_MENDSUM = 34 /* int64: sum of the machines' completion times */,
    FJ_S_INV_KT = 36 /* double 1/KT */, FJ_S_INV_M = 38 /* double 1/M: divisors of observation-only means */, FJ_S_COUNT = 40
};

// error flags (FJ_S_ERROR), same meaning as the oracle's
enum { FJ_E_LP = 1, FJ_E_NO_TASK = 2, FJ_E_NO_MACHINE = 4, FJ_E_NO_EVENT = 8, FJ_E_OVERFLOW = 16 };

struct FjParams {
    FjDims d;
    FjInstOff io;
    FjEnvOff eo;
    const int32_t *inst;        // instance table
    const int32_t *env_inst;    // [B] instance index of each env
    const int32_t *order;       // [n_slots] env of every warp slot of the main kernel (slot = virtual CTA * warps + warp), -1 = empty
    int n_slots;
    unsigned char *env;         // env table
    unsigned char *lp;          // LP scratch slabs (one per CTA of the main kernel / per warp of the resume kernel)
    unsigned long long lp_stride;
    int *pend_count;            // [FJ_ROUNDS + 1] parked LPs per resume round of the current launch; [FJ_ROUNDS + 1]: env CTAs finished
    int *pend_env;              // [2][B] env of each parked LP (ping-pong between rounds)
    double *lp_x;               // [lp_slots][NPx] LP solutions
    int *lp_meta;               // [lp_slots][2] iterations, return code
    int lp_slots;
    const double *plan_x;       // [n_instances][NPx] cached order-0 LP solution per instance
    const int *plan_meta;       // [n_instances][2]
    const int *plan_ok;         // [n_instances] 1 once cached (null before the first reset)
    int stage_stride;           // bytes of one warp's shared-memory slab (env hot prefix + instance hot words)
    int env_warps;              // env warps per CTA of the main kernel (= warp slots of a virtual CTA)
    int cta_lp;                 // 1: the LP-server CTAs of the main kernel solve the order-arrival LPs; 0: park for the LP / resume kernels; 2: free-running warps
    // LP service inside the main kernel (cta_lp == 1): its first `srv_ctas` CTAs run no environments; their warps
    // form `srv_groups` groups of `srv_group_warps` warps that serve the requests the env warps post on a queue in HBM/L2
    int srv_ctas, srv_groups, srv_group_warps, srv_group_smem;   // srv_group_smem: bytes of dynamic shared memory per group
    unsigned int *lpq;          // [0] head (claimed tickets), [1] tail (issued tickets): never reset
    unsigned long long *lpq_ring;   // [FJ_LPQ_RING] (ticket + 1) << 32 | env-warp slot
    int *lp_req;                // [env CTAs x env_warps][lp_req_stride]: env, then fstart[KTx], then the empty-queue mask [KTW + 1]
    int lp_req_stride;
    int *lp_resp;               // [env CTAs x env_warps][4]: ready flag, iterations, return code
    unsigned char *lp_own;      // one LP scratch slab (lp_stride bytes) per env warp of the launch for the overflow path, or null
    int lp_own_slots;           // env warps that have one
    int lp_overflow;            // queue depth from which an env warp solves its LP itself (0: never)
    int lock_groups;            // lockstep groups of an env CTA: 1, 2 (warps by pairs of SM sub-partitions) or 4 (one per sub-partition)
    int srv_join;               // an env CTA whose envs have all finished the launch serves LPs until the launch ends
    double *cta_x;              // [env CTAs][env_warps][NPx] LP solutions, one buffer per env warp
    int lock_mask;              // the lockstep group of an env CTA meets at a barrier every lock_mask + 1 steps (power of two - 1)
    int stage;                  // 1: kernels stage the hot part of the env record in shared memory
    int B, variant, sum_mode, nobs;
    long long *trace;           // FJ_TRACE builds: [grid][33][8] per-warp cycle counters + one row of LP phase cycles per CTA (null otherwise)
};

struct FjStepArgs {
    int T;                      // steps per launch
    const int32_t *actions;     // [T][B][2]
    const uint32_t *rnd;        // [T][B][2]
    int reward_policy, autoreset;
    double completion, tardiness, energy;
    double *state;              // [T][B][2*nobs] or null
    float *state32;             // [T][B][2*nobs] or null
    double *reward;             // [T][B] or null
    int32_t *done;              // [T][B] or null
    int32_t *rec;               // [T][B][8] or null
    int *park_count;            // where this kernel parks envs that need a fluid LP
    int *park_env;
    // progress of the launch for the host-buffer entry point (null: off).  Every env adds one to
    // prog_count[c] when its outputs of steps [c << prog_shift, (c + 1) << prog_shift) are written; the env that
    // completes a chunk stores prog_seq into prog_flag[c] (page-locked host memory): the host then copies that
    // chunk's outputs out with the copy engine while the kernel plays the next steps
    unsigned *prog_count = nullptr;
    unsigned *prog_flag = nullptr;
    unsigned prog_seq = 0;
    int prog_shift = 0;
};

#define FJ_LPQ_RING 8192   // power of two, larger than the env warps of a launch (each has at most one request outstanding)
#define FJ_ROUNDS 2   // resume rounds per launch; a third LP of one env inside a launch is solved in line

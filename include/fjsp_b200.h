/* fjsp_b200.h -- C ABI of the B200 batched FJSP scheduling environment.
 *
 * Drop-in boundary for the reference's per-instance Python environments
 *   environments/SO_DFJSP.py:13-268           SO_DFJSP_Environment.reset()/step(action)
 *   environments/MO_DFJSP.py:12-298           MO_DFJSP_Environment.reset()/step(action, reward_policy, ...)
 *   environments/MO_DFJSP_breakdown.py:12-328 same, with machine breakdown / repair intervals
 *   environments/SO_FJSSP.py:12-265           SO_FJSSP_Environment (per-job due dates of class_FJSSP.py)
 * and for the process-pool rollout of utilities/Parallel_Experience_Generator.py:28-66
 * (play_n_episodes / play_1_episode): one call steps EVERY environment copy of a batch on
 * the GPU.  Plain pointers and sizes only; a binding needs no CUDA or torch headers.
 *
 * Instance blob (int32 words), one per distinct problem instance:
 *   [0] 0x464A5350  [1] total words  [2] M machines  [3] K job kinds  [4] KT operation types
 *   [5] S orders  [6] NP eligible (machine, operation-type) pairs  [7] NBD breakdown intervals
 *   [8],[9] float64 bits (lo,hi) of the DDT state feature  [10] max jobs of one kind
 *   [11] total jobs  [12..15] 0
 *   ntask[K] rj_kind[KT] rj_stage[KT] nelig[KT] mt_order[KT*M] ptime[KT*M] power[KT*M]
 *   idle_power[M] arrive[S] due[S] count[S*K] bd_ptr[M+1] bd_start[NBD] bd_end[NBD]
 *   pair_order[NP]
 * (deep_reinforcement_learning_for_fjsp_b200/instance.py builds it from the reference's csv files or generators.)
 *
 * Every function returns 0 on success or a negative code; fjsp_last_error() describes
 * the failure.  There is no CPU fallback: without a CUDA device create() fails.
 */
#ifndef FJSP_B200_H
#define FJSP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fjsp_vec fjsp_vec;

enum { FJSP_VARIANT_SO_DFJSP = 0, FJSP_VARIANT_MO_DFJSP = 1, FJSP_VARIANT_MO_DFJSP_BREAKDOWN = 2,
       FJSP_VARIANT_SO_FJSSP = 3 };
/* sum_mode: how the reference's builtin sum() adds floats. 1 = CPython >= 3.12
 * (Neumaier-compensated), 0 = CPython <= 3.11 (left to right). */

const char *fjsp_last_error(void);
int fjsp_abi_version(void);

/* Builds the device tables.  blobs: all instance blobs back to back; blob_offsets[i]: word
 * offset of instance i; env_instance[e]: which instance environment e plays
 * (replaces constructing n_envs reference Environment objects). */
int fjsp_vec_create(const int32_t *blobs, const int64_t *blob_offsets, int n_instances,
                    const int32_t *env_instance, int n_envs, int variant, int sum_mode, int device,
                    fjsp_vec **out);
int fjsp_vec_destroy(fjsp_vec *v);

/* out[0..11] = n_envs, state_size (20 SO / 30 MO), env record bytes, instance record bytes,
 * grid blocks, threads per block, LP scratch bytes per slab, kernel launches so far,
 * env warps per block (= warp slots of a virtual CTA), LP-server blocks of the grid, warp slots,
 * dynamic shared memory bytes per block of the step kernel */
int fjsp_vec_query(fjsp_vec *v, int64_t *out12);

/* reset() of every environment (SO_DFJSP.py:54-79, MO_DFJSP.py:58-89).  d_state64 /
 * d_state32: DEVICE buffers [n_envs][state_size] (either may be null).  stream: a
 * cudaStream_t passed as void* (null = default stream). */
int fjsp_vec_reset(fjsp_vec *v, void *stream, double *d_state64, float *d_state32);

/* T consecutive step(action) calls of every environment in ONE launch (T = 1 is the
 * reference's step()).  All pointers are DEVICE pointers; outputs may be null.
 *   d_actions [T][n_envs][2]  (task rule, machine rule), 0-based as in the reference
 *   d_rnd     [T][n_envs][2]  32-bit draws consumed by the random rules (null = zeros)
 *   reward_policy, completion, tardiness, energy: MO_DFJSP.step keyword arguments
 *   autoreset: a finished environment is reset before its next action
 *   d_state64/32 [T][n_envs][state_size], d_reward [T][n_envs], d_done [T][n_envs],
 *   d_rec [T][n_envs][8] = operation type, kind, stage, job number, machine, begin, end,
 *   machine completion (the schedule the reference keeps in its Task objects). */
int fjsp_vec_step(fjsp_vec *v, void *stream, int T, const int32_t *d_actions, const uint32_t *d_rnd,
                  int reward_policy, double completion, double tardiness, double energy, int autoreset,
                  double *d_state64, float *d_state32, double *d_reward, int32_t *d_done, int32_t *d_rec);

/* Same with HOST buffers: copies actions/draws in, runs the launch, delivers the requested
 * outputs and synchronises.  This is the call a Python/ctypes agent loop makes.  Page-locked
 * (mapped) output buffers are stored by the kernel itself while it runs; other buffers are
 * copied out of device staging in chunks of steps while the kernel plays the following ones
 * (T >= 16), or after the launch. */
int fjsp_vec_step_host(fjsp_vec *v, int T, const int32_t *h_actions, const uint32_t *h_rnd,
                       int reward_policy, double completion, double tardiness, double energy, int autoreset,
                       double *h_state64, float *h_state32, double *h_reward, int32_t *h_done, int32_t *h_rec);
int fjsp_vec_reset_host(fjsp_vec *v, double *h_state64, float *h_state32);

/* Pipelined form of fjsp_vec_step_host (same arguments) for callers whose actions do not depend on
 * the previous call's outputs -- the `for worker in pool` of Parallel_Experience_Generator.play_n_episodes
 * (:28-40) with rule-based or replayed actions: _begin queues the input copy and the launch and returns,
 * _wait blocks until the OLDEST call begun has finished (its outputs are then complete).  At most two
 * calls in flight; the copies of one (each direction on its own copy engine) overlap the kernels of the
 * other.  Host buffers should be page-locked (pageable ones make the copies synchronous). */
int fjsp_vec_step_host_begin(fjsp_vec *v, int T, const int32_t *h_actions, const uint32_t *h_rnd,
                             int reward_policy, double completion, double tardiness, double energy, int autoreset,
                             double *h_state64, float *h_state32, double *h_reward, int32_t *h_done, int32_t *h_rec);
int fjsp_vec_step_host_wait(fjsp_vec *v);

/* Per environment 12 int64: step_time, step_count, completion_time, delay_time_sum,
 * energy_consumption, lp_solves, lp_iterations, error flags, done, next order, episodes,
 * delay_time_sum_unprocessed.  Host buffer [n_envs][12]. */
int fjsp_vec_info(fjsp_vec *v, int64_t *h_out);

/* Diagnostic: the warp-slot map of the last step launch (environment of every warp slot of the
 * step kernel, slot = virtual CTA * warps per CTA + warp, -1 = empty; written before every launch
 * by the LP-aware packing).  h_out == NULL returns the number of slots, else fills h_out
 * (capacity entries) and returns the number of slots; negative on error. */
int fjsp_vec_slots(fjsp_vec *v, int32_t *h_out, int capacity);

/* Diagnostic (builds with -DFJ_TRACE only, otherwise returns -6): per warp of the step
 * kernel's grid 8 int64 [grid][33][8] = cycles in the rollout, in dispatch, clock loop, CTA LP
 * service, observation/outputs, LPs served by the CTA, spare, SM id (row 32 of a CTA: cycles per
 * phase of its LPs).  clear != 0 zeroes them. */
int fjsp_vec_trace(fjsp_vec *v, int64_t *h_out, int clear);

#ifdef __cplusplus
}
#endif
#endif

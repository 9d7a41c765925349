/* TEST INFRASTRUCTURE ONLY (oracle/): pthread driver that steps many oracle
 * environments on the host cores.  Used by bench.py's cpu_baseline / --impl
 * reference legs (the "port" CPU baseline) and by tests that compare whole batches. */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <unistd.h>

int fjsp_oracle_step(void *h, int task_rule, int machine_rule, uint32_t rnd_task, uint32_t rnd_machine,
                     int reward_policy, double completion, double tardiness, double energy_norm,
                     double *state_out, double *reward_out, int *done_out, int *rec);
int fjsp_oracle_reset(void *h, double *state_out);
int fjsp_oracle_done(void *h);

int fjsp_oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    void **envs; int B, T; const int *actions; const uint32_t *rnd; int reward_policy, nstate;
    double *state, *reward; int *done, *rec; int *next; pthread_mutex_t *mu; int err;
} Job;

static void *worker(void *arg)
{
    Job *jb = (Job *)arg;
    for (;;) {
        pthread_mutex_lock(jb->mu);
        int b = (*jb->next)++;
        pthread_mutex_unlock(jb->mu);
        if (b >= jb->B) break;
        double st[64], rw = 0; int dn = 0; int rc[8];
        for (int t = 0; t < jb->T; ++t) {
            size_t i = (size_t)t * jb->B + b;
            if (fjsp_oracle_done(jb->envs[b])) jb->err |= fjsp_oracle_reset(jb->envs[b], st);
            jb->err |= fjsp_oracle_step(jb->envs[b], jb->actions[2 * i], jb->actions[2 * i + 1],
                                        jb->rnd[2 * i], jb->rnd[2 * i + 1], jb->reward_policy,
                                        1.0, 1.0, 1.0, st, &rw, &dn, rc);
            if (jb->state) for (int k = 0; k < jb->nstate; ++k) jb->state[i * jb->nstate + k] = st[k];
            if (jb->reward) jb->reward[i] = rw;
            if (jb->done) jb->done[i] = dn;
            if (jb->rec) for (int k = 0; k < 8; ++k) jb->rec[i * 8 + k] = rc[k];
        }
    }
    return NULL;
}

/* T steps for each of B environments; actions [T][B][2], rnd [T][B][2];
 * outputs state [T][B][nstate], reward [T][B], done [T][B], rec [T][B][8] (any may be NULL).
 * A finished environment is reset right before its next action (lazy auto-reset, like the device). */
int fjsp_oracle_batch_rollout(void **envs, int B, int T, const int *actions, const uint32_t *rnd,
                              int reward_policy, int nstate, double *state, double *reward,
                              int *done, int *rec, int threads)
{
    if (threads <= 0) threads = fjsp_oracle_max_threads();
    if (threads > 256) threads = 256;
    if (threads > B) threads = B;
    pthread_t tid[256];
    Job jobs[256];
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int next = 0, err = 0;
    if (threads == 1) {   /* one environment (or one core): no thread to start */
        Job j = { envs, B, T, actions, rnd, reward_policy, nstate, state, reward, done, rec, &next, &mu, 0 };
        worker(&j);
        return j.err;
    }
    for (int i = 0; i < threads; ++i) {
        Job j = { envs, B, T, actions, rnd, reward_policy, nstate, state, reward, done, rec, &next, &mu, 0 };
        jobs[i] = j;
        pthread_create(&tid[i], NULL, worker, &jobs[i]);
    }
    for (int i = 0; i < threads; ++i) { pthread_join(tid[i], NULL); err |= jobs[i].err; }
    return err;
}

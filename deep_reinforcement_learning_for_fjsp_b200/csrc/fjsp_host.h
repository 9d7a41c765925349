// Host-side table builder: instance blobs (see include/fjsp_b200.h) -> the instance
// table and layout the kernels consume.  Pure C++ (no CUDA), shared by the CUDA
// library (fjsp_api.cu) and by the host simulation of the kernel source used in
// the CPU tests (tests/hostsim/hostsim.cpp).
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "fjsp_layout.h"

struct FjBlobView {
    int M, K, KT, S, NP, NBD, Nmax, NJ;
    int ddt_lo, ddt_hi;
    const int32_t *ntask, *rj_kind, *rj_stage, *nelig, *mt_order, *ptime, *power, *idle_power;
    const int32_t *arrive, *due, *count, *bd_ptr, *bd_start, *bd_end, *pair_order;
};

static inline bool fj_parse_blob(const int32_t *b, FjBlobView &v, std::string &err)
{
    if (b[0] != FJSP_MAGIC) { err = "instance blob: bad magic"; return false; }
    v.M = b[2]; v.K = b[3]; v.KT = b[4]; v.S = b[5]; v.NP = b[6]; v.NBD = b[7];
    v.ddt_lo = b[8]; v.ddt_hi = b[9]; v.Nmax = b[10]; v.NJ = b[11];
    if (v.M < 1 || v.M > FJSP_MAX_M) { err = "instance blob: machine count outside 1..32"; return false; }
    if (v.KT < 1 || v.KT > FJSP_MAX_KT) { err = "instance blob: operation types outside 1..256"; return false; }
    if (v.S < 1 || v.S > FJSP_MAX_S) { err = "instance blob: order count outside 1..16"; return false; }
    if (v.Nmax > 65000) { err = "instance blob: more than 65000 jobs of one kind"; return false; }
    if (v.K < 1 || v.K > v.KT || v.K > 255) { err = "instance blob: job kinds outside 1..min(KT, 255)"; return false; }
    if (v.NP < v.KT || v.NP > v.KT * v.M || v.NBD < 0) { err = "instance blob: pair / breakdown counts out of range"; return false; }
    const int32_t *p = b + 16;
    v.ntask = p; p += v.K;
    v.rj_kind = p; p += v.KT;
    v.rj_stage = p; p += v.KT;
    v.nelig = p; p += v.KT;
    v.mt_order = p; p += v.KT * v.M;
    v.ptime = p; p += v.KT * v.M;
    v.power = p; p += v.KT * v.M;
    v.idle_power = p; p += v.M;
    v.arrive = p; p += v.S;
    v.due = p; p += v.S;
    v.count = p; p += v.S * v.K;
    v.bd_ptr = p; p += v.M + 1;
    v.bd_start = p; p += v.NBD;
    v.bd_end = p; p += v.NBD;
    v.pair_order = p; p += v.NP;
    if (p - b != b[1]) { err = "instance blob: length does not match header"; return false; }
    // every field the table builder indexes with (the blob comes through a public C ABI)
    for (int r = 0; r < v.K; ++r)
        if (v.ntask[r] < 1 || v.ntask[r] > 128) { err = "instance blob: operations per job kind outside 1..128"; return false; }
    for (int q = 0; q < v.KT; ++q) {
        if (v.rj_kind[q] < 0 || v.rj_kind[q] >= v.K || v.rj_stage[q] < 0 || v.rj_stage[q] >= v.ntask[v.rj_kind[q]]) {
            err = "instance blob: rj_kind / rj_stage out of range"; return false;
        }
        if (v.nelig[q] < 1 || v.nelig[q] > v.M) { err = "instance blob: eligible machines of an operation type outside 1..M"; return false; }
        unsigned seen = 0;
        for (int k = 0; k < v.nelig[q]; ++k) {
            const int m = v.mt_order[q * v.M + k];
            if (m < 0 || m >= v.M || (seen >> m & 1u)) { err = "instance blob: mt_order entry out of range or repeated"; return false; }
            seen |= 1u << m;
            if (v.ptime[q * v.M + m] < 1) { err = "instance blob: processing time of an eligible pair < 1"; return false; }
        }
    }
    for (int k = 0; k < v.NP; ++k)
        if (v.pair_order[k] < 0 || v.pair_order[k] >= v.KT * v.M) { err = "instance blob: pair_order entry out of range"; return false; }
    for (int m = 0; m <= v.M; ++m)
        if (v.bd_ptr[m] < 0 || v.bd_ptr[m] > v.NBD || (m && v.bd_ptr[m] < v.bd_ptr[m - 1])) { err = "instance blob: bd_ptr not a prefix array"; return false; }
    for (int i = 0; i < v.S * v.K; ++i)
        if (v.count[i] < 1 || v.count[i] > 65000) { err = "instance blob: per-order job count outside 1..65000"; return false; }
    return true;
}

// Iteration order of CPython's set(seq) for distinct ints in [0,32)
// (Objects/setobject.c: 8-slot table up to 4 entries, then 32 slots = ascending).
static inline int fj_pyset_order(const int *seq, int n, int *out)
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
        unsigned i = (unsigned)seq[e] & 7u, perturb = (unsigned)seq[e];
        while (slot[i] >= 0) { perturb >>= 5; i = (i * 5u + 1u + perturb) & 7u; }
        slot[i] = seq[e];
    }
    int k = 0;
    for (int i = 0; i < 8; ++i) if (slot[i] >= 0) out[k++] = slot[i];
    return k;
}

static inline int fj_align(int x, int a) { return (x + a - 1) / a * a; }

struct FjTables {
    FjDims d;
    FjInstOff io;
    FjEnvOff eo;
    std::vector<int32_t> inst;   // n_instances * io.stride
    int n_instances;
};

static inline bool fj_build_tables(const int32_t *blobs, const int64_t *offsets, int n_inst, FjTables &t,
                                   std::string &err, int variant = 0)
{
    std::vector<FjBlobView> views(n_inst);
    FjDims d;
    memset(&d, 0, sizeof(d));
    for (int i = 0; i < n_inst; ++i) {
        if (!fj_parse_blob(blobs + offsets[i], views[i], err)) return false;
        const FjBlobView &v = views[i];
        if (v.M > d.Mx) d.Mx = v.M;
        if (v.K > d.Kx) d.Kx = v.K;
        if (v.KT > d.KTx) d.KTx = v.KT;
        if (v.S > d.Sx) d.Sx = v.S;
        if (v.NBD > d.NBDx) d.NBDx = v.NBD;
        if (v.NJ > d.NJx) d.NJx = v.NJ;
        if (v.NP > d.NPx) d.NPx = v.NP;
        if (variant == FJSP_SO_FJSSP && (v.Nmax + 31) / 32 > d.NWx) d.NWx = (v.Nmax + 31) / 32;
    }
    d.KTW = (d.KTx + 31) / 32;
    d.Rx = d.Mx + 2 * d.KTx;
    d.NFx = d.Rx;
    t.d = d;
    // instance offsets (words)
    FjInstOff io;
    int o = 0;
    // hot head: small arrays every step reads (staged in shared memory by the main kernel)
    io.hdr = o; o += 12;
    io.ntask = o; o += d.Kx;
    io.first = o; o += d.Kx;
    io.jobbase = o; o += d.Kx;
    io.rjkind = o; o += d.KTx;
    io.rjstage = o; o += d.KTx;
    io.rjlast = o; o += d.KTx;
    io.elig = o; o += d.KTx;
    io.nelig = o; o += d.KTx;
    io.colbase = o; o += d.KTx;
    io.mnkt = o; o += d.Mx;
    io.idlep = o; o += d.Mx;
    io.arrive = o; o += d.Sx;
    io.due = o; o += d.Sx;
    io.count = o; o += d.Sx * d.Kx;
    io.cum = o; o += (d.Sx + 1) * d.Kx;
    o = fj_align(o, 4);
    io.hotw = o;
    // per-pair tables and breakdown lists: read through ld.global.nc
    io.mtset = o; o += d.KTx * d.Mx;
    io.mtpack = o; o += d.KTx;
    io.poord = o; o += d.KTx * d.Mx;
    io.ptime = o; o += d.KTx * d.Mx;
    io.energy = o; o += d.KTx * d.Mx;
    io.bdptr = o; o += d.Mx + 1;
    io.bds = o; o += d.NBDx;
    io.bde = o; o += d.NBDx;
    io.colqm = o; o += d.NPx;
    o = fj_align(o, 2);
    io.colrate = o; o += 2 * d.NPx;
    io.stride = fj_align(o, 4);
    t.io = io;
    // env offsets (bytes)
    FjEnvOff eo;
    int b = 0;
    // hot prefix: what every step reads (staged in shared memory by the main kernel)
    eo.scal = b; b += 4 * FJ_S_COUNT;
    eo.obs = b; b += 8 * 16;
    eo.obs2 = b; b += 8 * 16;
    eo.gapave = b; b += 8 * d.Mx;
    eo.mF = b; b += 8 * d.Mx;
    eo.invnkt = b; b += 8 * d.Mx;
    eo.rsum = b; b += 8 * d.KTx;
    eo.tsum = b; b += 8 * d.KTx;
    eo.avmask = b; b += 4 * d.KTW;
    eo.favmask = b; b += 4 * d.KTW;
    eo.demask = b; b += 4 * d.KTW;
    eo.damask = b; b += 4 * d.KTW;
    eo.mend = b; b += 4 * d.Mx;
    eo.mlast = b; b += 4 * d.Mx;
    eo.mjob = b; b += 4 * d.Mx;
    eo.mD = b; b += 4 * d.Mx;
    eo.proc = b; b += 4 * d.KTx;
    eo.fstart = b; b += 4 * d.KTx;
    eo.flmask = b; b += 4 * d.KTx;
    eo.qhead = b; b += 2 * d.KTx;
    eo.qtail = b; b += 2 * d.KTx;
    eo.qlen = b; b += 2 * d.KTx;
    eo.cntunp = b; b += 2 * d.KTx * d.Sx;
    eo.cntnow = b; b += 2 * d.KTx * d.Sx;
    b = fj_align(b, 4);
    eo.h_elig = b; b += 4 * d.KTx;
    eo.h_mtpack = b; b += 4 * d.KTx;
    eo.h_fo = b; b += 4 * d.KTx;
    eo.h_due = b; b += 4 * d.Sx;
    eo.h_cum = b; b += 4 * (d.Sx + 1) * d.Kx;
    eo.h_jobbase = b; b += 4 * d.Kx;
    eo.h_rjinfo = b; b += 2 * d.KTx;
    b = fj_align(b, 16);
    eo.hot = b;
    // the rest stays in HBM/L2: rule keys of available types, per-pair counters, fluid slots, links
    eo.urg = b; b += 8 * d.KTx;
    eo.maxe = b; b += 8 * d.KTx;
    eo.fu = b; b += 8 * d.NFx;
    eo.fa = b; b += 8 * d.NFx;
    eo.ff = b; b += 8 * d.NFx;
    eo.pk = b; b += 2 * d.KTx * d.Mx;
    eo.slot = b; b += 2 * d.KTx * d.Mx;
    eo.next = b; b += 2 * d.NJx;
    b = fj_align(b, 4);
    eo.unpmask = b; b += 4 * d.KTx * d.NWx;
    eo.duejob = b; b += (d.NWx ? 4 * d.NJx : 0);
    eo.mindue = b; b += (d.NWx ? 4 * d.KTx : 0);
    eo.stride = fj_align(b, 16);
    t.eo = eo;
    // fill instance records
    t.n_instances = n_inst;
    t.inst.assign((size_t)n_inst * io.stride, 0);
    for (int i = 0; i < n_inst; ++i) {
        const FjBlobView &v = views[i];
        int32_t *w = t.inst.data() + (size_t)i * io.stride;
        int32_t *h = w + io.hdr;
        h[0] = v.M; h[1] = v.K; h[2] = v.KT; h[3] = v.S; h[4] = v.NJ; h[5] = v.ddt_lo; h[6] = v.ddt_hi; h[7] = v.NP;
        {   // h[8]: operations of one episode (every step dispatches one)
            long long ops = 0;
            for (int r = 0; r < v.K; ++r) { long long c = 0; for (int s = 0; s < v.S; ++s) c += v.count[s * v.K + r]; ops += c * v.ntask[r]; }
            h[8] = (int32_t)(ops > 0x7fffffff ? 0x7fffffff : ops);
        }
        int acc = 0, jb = 0;
        for (int r = 0; r < v.K; ++r) {
            w[io.ntask + r] = v.ntask[r];
            w[io.first + r] = acc; acc += v.ntask[r];
            w[io.jobbase + r] = jb;
            int c = 0;
            for (int s = 0; s < v.S; ++s) {
                w[io.cum + s * d.Kx + r] = c;
                w[io.count + s * d.Kx + r] = v.count[s * v.K + r];
                c += v.count[s * v.K + r];
            }
            for (int s = v.S; s <= d.Sx; ++s) w[io.cum + s * d.Kx + r] = c;
            jb += c;
        }
        if (acc != v.KT) { err = "instance blob: sum(ntask) != KT"; return false; }
        // rank of every pair in x.items() order
        std::vector<int> rank((size_t)v.KT * v.M, -1);
        for (int k = 0; k < v.NP; ++k) rank[v.pair_order[k]] = k;
        for (int q = 0; q < v.KT; ++q) {
            int r = v.rj_kind[q], j = v.rj_stage[q];
            w[io.rjkind + q] = r; w[io.rjstage + q] = j;
            w[io.rjlast + q] = (j == v.ntask[r] - 1);
            w[io.nelig + q] = v.nelig[q];
            unsigned mask = 0;
            int seq[32], ord[32];
            for (int k = 0; k < v.nelig[q]; ++k) { seq[k] = v.mt_order[q * v.M + k]; mask |= 1u << seq[k]; }
            w[io.elig + q] = (int32_t)mask;
            int n = fj_pyset_order(seq, v.nelig[q], ord);
            for (int k = 0; k < d.Mx; ++k) w[io.mtset + q * d.Mx + k] = k < n ? ord[k] : -1;
            {
                unsigned pk = 0;
                if (n <= 4) for (int k = 0; k < n; ++k) pk |= (unsigned)ord[k] << (8 * k);
                w[io.mtpack + q] = (int32_t)pk;
            }
            // machines of q sorted by pair rank
            int cnt = 0;
            for (int m = 0; m < v.M; ++m) if (mask >> m & 1) {
                if (rank[q * v.M + m] < 0) { err = "instance blob: pair_order misses an eligible pair"; return false; }
                int pos = cnt++;
                while (pos > 0 && rank[q * v.M + w[io.poord + q * d.Mx + pos - 1]] > rank[q * v.M + m]) {
                    w[io.poord + q * d.Mx + pos] = w[io.poord + q * d.Mx + pos - 1];
                    --pos;
                }
                w[io.poord + q * d.Mx + pos] = m;
            }
            for (int k = cnt; k < d.Mx; ++k) w[io.poord + q * d.Mx + k] = -1;
            for (int m = 0; m < v.M; ++m) {
                w[io.ptime + q * d.Mx + m] = v.ptime[q * v.M + m];
                w[io.energy + q * d.Mx + m] = v.power[q * v.M + m] * v.ptime[q * v.M + m];
            }
        }
        {
            int cb = 0;
            for (int q = 0; q < v.KT; ++q) {
                w[io.colbase + q] = cb;
                // LP columns of the type: machine ascending; 1.0 / p is the IEEE quotient the kernels used to divide for
                for (int m = 0; m < v.M; ++m) if ((unsigned)w[io.elig + q] >> m & 1u) {
                    if (cb >= v.NP) { err = "instance blob: NP != sum(nelig)"; return false; }
                    w[io.colqm + cb] = (q << 8) | m;
                    const double rate = 1.0 / (double)v.ptime[q * v.M + m];
                    memcpy(&w[io.colrate + 2 * cb], &rate, 8);
                    ++cb;
                }
            }
            if (cb != v.NP) { err = "instance blob: NP != sum(nelig)"; return false; }
        }
        for (int m = 0; m < v.M; ++m) {
            w[io.idlep + m] = v.idle_power[m];
            int n = 0;
            for (int q = 0; q < v.KT; ++q) n += v.ptime[q * v.M + m] > 0;
            w[io.mnkt + m] = n;
        }
        for (int s = 0; s < v.S; ++s) { w[io.arrive + s] = v.arrive[s]; w[io.due + s] = v.due[s]; }
        for (int m = 0; m <= d.Mx; ++m) w[io.bdptr + m] = v.bd_ptr[m <= v.M ? m : v.M];
        for (int k = 0; k < v.NBD; ++k) { w[io.bds + k] = v.bd_start[k]; w[io.bde + k] = v.bd_end[k]; }
    }
    return true;
}

// bytes of LP scratch per resident warp: Binv, the small arrays, the solution vector
// (see fjsp_core.cuh: fj_lp_carve / fj_order_arrives_inline)
static inline unsigned long long fj_lp_small_bytes_host(const FjDims &d)
{
    unsigned long long R = d.Rx, C = d.NPx + 1;
    return (2 * R + 2 * C) * 8 + (R + (C + R) + 2 * C + d.KTx) * 4;
}
static inline unsigned long long fj_lp_scratch_bytes(const FjDims &d)
{
    unsigned long long R = d.Rx;
    unsigned long long bytes = R * R * 8 + (fj_lp_small_bytes_host(d) + 7) / 8 * 8 + (unsigned long long)d.NPx * 8;
    return (bytes + 127) / 128 * 128;
}

// sched_host.cpp -- host mirror of the tile-DAG scheduler in progress-word form.
//
// The reference keeps one Task slot per tile whose (k, taskStatus) pair encodes how far
// the tile has got (src/gridscheduler.c:51-64 genericdone).  Here that pair is collapsed
// into one integer per tile, prog[l][m] = number of completed steps, which is exactly the
// word the persistent CUDA kernel spins on.  Task (l,m,k) of type
//   QRS  (l==k, m==k)  needs prog[k][k]   == k
//   SAPP (l==k, m> k)  needs prog[k][m]   == k  and prog[k][k]   >= k+1
//   QRD  (l> k, m==k)  needs prog[l][k]   == k  and prog[l-1][k] >= k+1
//   DAPP (l> k, m> k)  needs prog[l][m]   == k  and prog[l][k]   >= k+1 and prog[l-1][m] >= k+1
// (gridscheduler.c:112-174 candoQRS/SAPP/QRD/DAPP).  Driven by one serial worker with the
// reference's pick rule (gridscheduler.c:259-300) it reproduces the reference's task
// sequence; tests/test_abi.py checks that against the golden traces.
#include "../../include/qrb200.h"

#include <cstdlib>
#include <vector>

namespace {
enum : unsigned char { S_IDLE = 0, S_READY = 1, S_DOING = 2 };

inline int task_type(int l, int m, int k)
{
    if (l == k) return m == k ? B200QR_QRS : B200QR_SAPP;
    return m == k ? B200QR_QRD : B200QR_DAPP;
}
}  // namespace

struct b200qr_sched_s {
    int M, N;
    std::vector<int> prog;              // column-major like the reference's tgrid: [m*M + l]
    std::vector<unsigned char> state;
    int ready;

    int &P(int l, int m) { return prog[(size_t)m * M + l]; }
    unsigned char &S(int l, int m) { return state[(size_t)m * M + l]; }
    bool inside(int l, int m) const { return l >= 0 && m >= 0 && l < M && m < N; }

    // all inputs of the tile's next task are final?
    bool runnable(int l, int m)
    {
        if (!inside(l, m)) return false;
        int k = P(l, m);
        if (k > (l < m ? l : m)) return false;          // tile has no task left
        if (l > k && P(l - 1, m) < k + 1) return false;  // tile above finished step k
        if (m > k && P(l, k) < k + 1) return false;      // panel tile of this row finished step k
        return true;
    }
    void promote(int l, int m)
    {
        if (inside(l, m) && S(l, m) == S_IDLE && runnable(l, m)) { S(l, m) = S_READY; ++ready; }
    }
};

extern "C" {

b200qr_sched_t b200qr_sched_init(int M, int N)
{
    if (M <= 0 || N <= 0) return nullptr;
    b200qr_sched_s *s = new b200qr_sched_s;
    s->M = M; s->N = N; s->ready = 0;
    s->prog.assign((size_t)M * N, 0);
    s->state.assign((size_t)M * N, S_IDLE);
    s->promote(0, 0);                                    // gridscheduler.c:320 seeds QRS(0,0)
    return s;
}

void b200qr_sched_free(b200qr_sched_t s) { delete s; }

int b200qr_sched_next(b200qr_sched_t s, b200qr_task *out)
{
    bool busy = false;
    for (int l = s->M - 1; l >= 0; --l)
        for (int m = s->N - 1; m >= 0; --m) {
            unsigned char st = s->S(l, m);
            if (st == S_READY) {
                s->S(l, m) = S_DOING; --s->ready;
                int k = s->P(l, m);
                out->type = task_type(l, m, k); out->l = l; out->m = m; out->k = k;
                return B200QR_TASK_AVAIL;
            }
            if (st == S_DOING) busy = true;
        }
    return busy ? B200QR_TASK_NONE : B200QR_TASK_DONE;
}

void b200qr_sched_done(b200qr_sched_t s, const b200qr_task *t)
{
    int l = t->l, m = t->m;
    int k = s->P(l, m);
    int type = task_type(l, m, k);
    s->S(l, m) = S_IDLE;
    s->P(l, m) = k + 1;
    // Who can this completion unblock?  Same neighbourhoods the reference probes
    // (gridscheduler.c:188-253), expressed on progress words.
    if (type == B200QR_QRS || type == B200QR_QRD)
        for (int j = k + 1; j < s->N; ++j) s->promote(l, j);   // the panel row to the right
    if (type == B200QR_DAPP) s->promote(l, m);                 // same tile, next step
    s->promote(l + 1, m);                                      // the tile below
}

int b200qr_sched_progress(b200qr_sched_t s, int l, int m) { return s->inside(l, m) ? s->P(l, m) : -1; }
int b200qr_sched_ready_count(b200qr_sched_t s) { return s->ready; }

// Value of the device's progress word of tile (l, m) once its tile column is final: the number of steps the tile takes
// part in, min(l, m) + 1 capped at K = min(p, q) -- what b200qr_sched_progress() reads when the DAG has run dry -- plus
// one on the diagonal, where the panel's last store publishes "column final" (run_panel in qr_persistent.cu).  The
// overlapped host path of cudaQRTask compares the words it polls with this.
int b200qr_sched_final_progress(int p, int q, int l, int m)
{
    if (p <= 0 || q <= 0 || l < 0 || m < 0 || l >= p || m >= q) return -1;
    const int K = p < q ? p : q, lo = l < m ? l : m;
    const int steps = lo + 1 < K ? lo + 1 : K;
    return steps + ((l == m && m < K) ? 1 : 0);
}

int b200qr_sched_dep_count(int type, int k)
{
    int prev = k > 0 ? 1 : 0;            // DAPP(l,m,k-1) on the same tile
    switch (type) {
    case B200QR_QRS:  return prev;
    case B200QR_SAPP: return 1 + prev;   // QRS(k,k,k)
    case B200QR_QRD:  return 1 + prev;   // task above in the panel column
    case B200QR_DAPP: return 2 + prev;   // QRD(l,k,k) and the task above in column m
    }
    return -1;
}

int b200qr_sched_predecessors(const b200qr_task *t, b200qr_task pred[3])
{
    int n = 0, l = t->l, m = t->m, k = t->k;
    if (k > 0) pred[n++] = b200qr_task{B200QR_DAPP, l, m, k - 1};
    if (m > k) pred[n++] = b200qr_task{task_type(l, k, k), l, k, k};        // QRS (l==k) or QRD
    if (l > k) pred[n++] = b200qr_task{task_type(l - 1, m, k), l - 1, m, k};
    return n;
}

// Sweep ticket order of the persistent kernel.  A sweep S(k,g) = UNMQR(k,j) + the TSMQR chain
// (i,j), i>k, for the tile columns j of group g (G columns per group).  Tickets are sorted by
//     key(k,g) = k + G*g      (ties: smaller k first)
// i.e. a group that is d groups to the right of the diagonal group of step k is handed out
// about d*G steps later than the diagonal group of the same step.  By then the panel chain
// P(k) has a long head start, so only the diagonal-group sweeps run in lock-step with a panel;
// everything else runs at full speed.  For a fixed group the key grows with k, and the panel
// P(k) only needs S(k-1, diagonal group), whose key is smaller than that of every S(k, .):
// the order is a topological order of the macro-task DAG (needed for deadlock freedom).
// out2 receives (k, g) pairs; returns the number of sweeps.
long b200qr_sched_sweep_order(int M, int N, int G, int *out2, long cap)
{
    return b200qr_sched_sweep_order_skew(M, N, G, G, out2, cap);
}

// Same with an explicit skew B: key(k,g) = k + B*g.  Any B >= 0 keeps the order topological (for a
// fixed group the key grows with k, and S(k-1, diagonal group) precedes every S(k, .)); a larger B
// hands the groups right of the diagonal out later, i.e. gives their panel a longer head start --
// needed when a sweep row is much faster than a TSQRT of the panel chain (tensor-core sweeps).
long b200qr_sched_sweep_order_skew(int M, int N, int G, int B, int *out2, long cap)
{
    if (M <= 0 || N <= 0 || G <= 0 || B < 0) return -1;
    const int K = M < N ? M : N;
    const int GB = (N + G - 1) / G;
    // counting sort on key in [0, K + G*GB]
    std::vector<long> count((size_t)K + (size_t)B * GB + 2, 0);
    long n = 0;
    for (int k = 0; k < K; ++k) {
        if (k + 1 >= N) continue;
        for (int g = (k + 1) / G; g < GB; ++g) { ++count[(size_t)k + (size_t)B * g + 1]; ++n; }
    }
    for (size_t i = 1; i < count.size(); ++i) count[i] += count[i - 1];
    for (int k = 0; k < K; ++k) {                       // k ascending => ties keep smaller k first
        if (k + 1 >= N) continue;
        for (int g = (k + 1) / G; g < GB; ++g) {
            long pos = count[(size_t)k + (size_t)B * g]++;
            if (pos < cap) { out2[2 * pos] = k; out2[2 * pos + 1] = g; }
        }
    }
    return n;
}

// Lagged order (tensor-core engine 3).  With four tile columns per sweep the reflector / T operands of a tile row
// are 1.5x the bytes of the row itself, so sweeps that run at the same time should work on the SAME step k: then the
// step's operand tiles are read from HBM once and served from L2 to everyone else.  The skewed order above does the
// opposite (at any time every group is at a different step).  Here a sweep is
//     far     if its group does not become the diagonal group within the next LAG steps:  key = k + LAG
//             -- all far sweeps of a step are adjacent in the order, LAG steps behind its panel (which is done by then);
//     urgent  otherwise:                                                                    key = max(k, kf + LAG)
//             -- in step order, right behind the panels that need them; kf = G*g - LAG - 2 is the group's last far step,
//             so the group's own history stays in order (ties keep smaller k first).
// For a fixed group the key never decreases with k, and P(k) needs S(k-1, diagonal group of k-1) whose key is k-1:
// a topological order of the macro-task DAG.
long b200qr_sched_sweep_order_lag(int M, int N, int G, int LAG, int *out2, long cap)
{
    if (M <= 0 || N <= 0 || G <= 0 || LAG < 0) return -1;
    const int K = M < N ? M : N;
    const int GB = (N + G - 1) / G;
    auto key = [&](int k, int g) -> long {
        if (g > (k + LAG + 1) / G) return (long)k + LAG;
        const long kf = (long)G * g - LAG - 2;
        return kf >= 0 && kf + LAG > k ? kf + LAG : k;
    };
    std::vector<long> count((size_t)K + (size_t)LAG + (size_t)G * GB + 4, 0);
    long n = 0;
    for (int k = 0; k < K; ++k) {
        if (k + 1 >= N) continue;
        for (int g = (k + 1) / G; g < GB; ++g) { ++count[(size_t)key(k, g) + 1]; ++n; }
    }
    for (size_t i = 1; i < count.size(); ++i) count[i] += count[i - 1];
    for (int k = 0; k < K; ++k) {                       // k ascending => ties keep smaller k first
        if (k + 1 >= N) continue;
        for (int g = (k + 1) / G; g < GB; ++g) {
            long pos = count[(size_t)key(k, g)]++;
            if (pos < cap) { out2[2 * pos] = k; out2[2 * pos + 1] = g; }
        }
    }
    return n;
}

// Tickets of one rank of the 1-D block-cyclic distribution (block = G tile columns, block g on
// rank g % world): the sweeps of its blocks, in the global skewed order, and the panels of its
// tile columns, ascending.  Every rank's streams are subsequences of the global streams, so the
// deadlock-freedom argument carries over unchanged.  Returns the sweep count; *npanels gets the
// panel count (either output array may be NULL to just count).
long b200qr_sched_rank_tickets(int M, int N, int G, int rank, int world, int *sweeps2, long cap, int *panels, long capp,
                               long *npanels)
{
    return b200qr_sched_rank_tickets_skew(M, N, G, G, rank, world, sweeps2, cap, panels, capp, npanels);
}

long b200qr_sched_rank_tickets_skew(int M, int N, int G, int B, int rank, int world, int *sweeps2, long cap, int *panels,
                                    long capp, long *npanels)
{
    if (world < 1 || rank < 0 || rank >= world) return -1;
    long total = b200qr_sched_sweep_order_skew(M, N, G, B, nullptr, 0);
    if (total < 0) return -1;
    std::vector<int> all(2 * (size_t)(total > 0 ? total : 1));
    b200qr_sched_sweep_order_skew(M, N, G, B, all.data(), total);
    long n = 0;
    for (long t = 0; t < total; ++t)
        if (all[2 * t + 1] % world == rank) {
            if (sweeps2 && n < cap) { sweeps2[2 * n] = all[2 * t]; sweeps2[2 * n + 1] = all[2 * t + 1]; }
            ++n;
        }
    const int K = M < N ? M : N;
    long np = 0;
    for (int k = 0; k < K; ++k)
        if ((k / G) % world == rank) {
            if (panels && np < capp) panels[np] = k;
            ++np;
        }
    if (npanels) *npanels = np;
    return n;
}

// One serial order that is consistent with both ticket streams of the persistent kernel (panel
// tickets P(0), P(1), ...; sweep tickets as above), expanded to tile tasks (n x 4 ints: type,l,m,k):
// each panel is emitted just before the first sweep that needs it.  Tests check that this is a
// topological order of the tile DAG.
long b200qr_sched_device_order(int M, int N, int *out4, long cap)
{
    const int G = 16;                                    // tile columns per sweep (2-column variant)
    int K = M < N ? M : N;
    long n = 0;
    auto emit = [&](int type, int l, int m, int k) {
        if (n < cap) { out4[4 * n] = type; out4[4 * n + 1] = l; out4[4 * n + 2] = m; out4[4 * n + 3] = k; }
        ++n;
    };
    auto panel = [&](int k) {
        emit(B200QR_QRS, k, k, k);
        for (int i = k + 1; i < M; ++i) emit(B200QR_QRD, i, k, k);
    };
    auto sweep = [&](int k, int g) {
        int j0 = g * G > k + 1 ? g * G : k + 1;
        int j1 = (g + 1) * G < N ? (g + 1) * G : N;
        for (int j = j0; j < j1; ++j) emit(B200QR_SAPP, k, j, k);
        for (int i = k + 1; i < M; ++i)                  // the consumer warps advance row by row together
            for (int j = j0; j < j1; ++j) emit(B200QR_DAPP, i, j, k);
    };
    long ns = b200qr_sched_sweep_order(M, N, G, nullptr, 0);
    std::vector<int> tab(2 * (size_t)(ns > 0 ? ns : 1));
    b200qr_sched_sweep_order(M, N, G, tab.data(), ns);
    int next_panel = 0;
    for (long t = 0; t < ns; ++t) {
        int k = tab[2 * t], g = tab[2 * t + 1];
        while (next_panel <= k) panel(next_panel++);
        sweep(k, g);
    }
    while (next_panel < K) panel(next_panel++);
    return n;
}

}  // extern "C"

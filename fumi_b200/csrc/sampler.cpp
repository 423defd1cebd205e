// Native episodic task sampler (host side).
//
// Bit-exact restatement of the reference loader for the path
//   BatchMetaDataLoader(ClassSplitter(InatAnim(...), shuffle=True, K, Q).seed(0), shuffle=True)
// (fumi/dataset/data.py:73-84,125-188; torchmeta 1.7.0 CombinationRandomSampler / ClassSplitter_ /
// Categorical).  Three generator streams decide a meta-batch:
//   1. CPython `random` (MT19937)  : class tuples, random.sample(range(C), N)
//   2. numpy RandomState           : per (task, class) hash-seeded permutation of the class's
//                                    images + the split's shared RandomState(0) shuffles
//   3. torch CPU generator         : label permutation torch.randperm(N); DataLoader base seed
// Stream 1 and 3 are owned by the host program (state passed in and written back); the hash-seeded
// generators are independent per (task, class) and are expanded on a thread pool.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fumi_b200.h"

void fumi_set_error(const std::string& msg);  // common.cpp

namespace {

struct MT19937 {                       // reference mt19937ar; shared by CPython, numpy and torch
    uint32_t key[624];
    uint32_t out[624];                 // tempered words of the current block (one vectorisable pass per 624 draws)
    int pos;                           // 624 => regenerate before the next draw
    void retemper() {                  // after key[] was written from outside (state import)
        for (int i = 0; i < 624; ++i) out[i] = temper(key[i]);
    }
    void seed(uint32_t s) {            // init_genrand
        for (int i = 0; i < 624; ++i) {
            key[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + uint32_t(i) + 1u;
        }
        pos = 624;
    }
    void twist() {
        const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
        int k = 0;
        for (; k < 624 - 397; ++k) {
            uint32_t y = (key[k] & UP) | (key[k + 1] & LO);
            key[k] = key[k + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        for (; k < 623; ++k) {
            uint32_t y = (key[k] & UP) | (key[k + 1] & LO);
            key[k] = key[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        uint32_t y = (key[623] & UP) | (key[0] & LO);
        key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        pos = 0;
        retemper();
    }
    static inline uint32_t temper(uint32_t y) {
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline uint32_t next() {
        if (pos >= 624) twist();
        return out[pos++];
    }
};

// numpy legacy random_interval(max): masked rejection sampling on 32-bit draws.
inline uint32_t np_interval(MT19937& g, uint32_t max) {
    if (max == 0) return 0;
    const uint32_t mask = 0xFFFFFFFFu >> __builtin_clz(max);       // smallest 2^k - 1 >= max  (max >= 1 here)
    uint32_t v;
    do { v = g.next() & mask; } while (v > max);
    return v;
}
// numpy RandomState.shuffle on a 1-D array: backward Fisher-Yates.
inline void np_shuffle(MT19937& g, int64_t* x, int64_t n) {
    for (int64_t i = n - 1; i >= 1; --i) {
        uint32_t j = np_interval(g, uint32_t(i));
        std::swap(x[i], x[j]);
    }
}

// CPython Random._randbelow_with_getrandbits.
inline uint32_t py_randbelow(MT19937& g, uint32_t n) {
    int k = 32 - __builtin_clz(n);       // n.bit_length(), n >= 1
    uint32_t r;
    do { r = g.next() >> (32 - k); } while (r >= n);
    return r;
}
// CPython random.sample(range(n), k) (Lib/random.py, 3.8+: pool path / set-rejection path).
void py_sample_range(MT19937& g, int64_t n, int k, int64_t* out) {
    int64_t setsize = 21;
    if (k > 5) {
        // 4 ** ceil(log(3k, 4))
        int64_t p = 1;
        while (p < int64_t(3) * k) p *= 4;
        setsize += p;
    }
    if (n <= setsize) {
        std::vector<int64_t> pool(n);
        for (int64_t i = 0; i < n; ++i) pool[i] = i;
        for (int i = 0; i < k; ++i) {
            uint32_t j = py_randbelow(g, uint32_t(n - i));
            out[i] = pool[j];
            pool[j] = pool[n - i - 1];
        }
    } else {
        for (int i = 0; i < k; ++i) {
            int64_t j;
            bool dup;
            do {
                j = py_randbelow(g, uint32_t(n));
                dup = false;
                for (int t = 0; t < i; ++t) dup |= (out[t] == j);
            } while (dup);
            out[i] = j;
        }
    }
}

// torch CPU generator (at::mt19937): same recurrence, but the regeneration test is on `left`.
struct TorchMT {
    uint32_t* st;  // [626]: key[624], next, left
    inline uint32_t next() {
        uint32_t& nxt = st[624];
        uint32_t& left = st[625];
        if (--left == 0) {
            MT19937 tmp;
            std::memcpy(tmp.key, st, sizeof(tmp.key));
            tmp.twist();
            std::memcpy(st, tmp.key, sizeof(tmp.key));
            left = 624;
            nxt = 0;
        }
        return MT19937::temper(st[nxt++]);
    }
};

}  // namespace

extern "C" int64_t fumi_py_tuple_hash(const int64_t* items, int64_t n) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t lane = uint64_t(items[i]);   // hash(int) == int for 0 <= int < 2^61-1
        acc += lane * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += uint64_t(n) ^ (P5 ^ 3527539ULL);
    if (acc == ~uint64_t(0)) return 1546275796;
    return int64_t(acc);
}

struct fumi_sampler {
    std::vector<int64_t> offsets, ids;
    int64_t C;
    int N, K, Q;
    MT19937 shared;  // ClassSplitter_.np_random = RandomState(0)
};

extern "C" int fumi_sampler_create(const int64_t* class_offsets, const int64_t* class_image_ids, int64_t C,
                                   int32_t N, int32_t K, int32_t Q, fumi_sampler** out) {
    if (!class_offsets || !class_image_ids || !out || C <= 0 || N <= 0 || K <= 0 || Q < 0 || N > C) {
        fumi_set_error("fumi_sampler_create: bad argument (need 0 < N <= C, K > 0, Q >= 0)");
        return FUMI_ERR_ARG;
    }
    fumi_sampler* s = new fumi_sampler();
    s->C = C; s->N = N; s->K = K; s->Q = Q;
    s->offsets.assign(class_offsets, class_offsets + C + 1);
    s->ids.assign(class_image_ids, class_image_ids + class_offsets[C]);
    s->shared.seed(0);
    *out = s;
    return FUMI_OK;
}

extern "C" void fumi_sampler_destroy(fumi_sampler* s) { delete s; }

extern "C" int fumi_sampler_new_iterator(fumi_sampler* s, uint32_t* torch_state) {
    if (!s || !torch_state) { fumi_set_error("fumi_sampler_new_iterator: null argument"); return FUMI_ERR_ARG; }
    TorchMT t{torch_state};
    t.next();  // torch.empty((), dtype=int64).random_(): one 64-bit draw = two u32
    t.next();
    return FUMI_OK;
}

extern "C" int fumi_sampler_next(fumi_sampler* s, int64_t B, uint32_t* py_state, uint32_t* torch_state,
                                 int64_t* classes, int64_t* label_perm, int64_t* sup_ids, int64_t* qry_ids,
                                 int64_t* sup_y, int64_t* qry_y, int64_t* head_class,
                                 int64_t* sup_rows, int64_t* qry_rows, int32_t num_threads) {
    if (!s || B <= 0 || !py_state || !torch_state || !classes || !label_perm || !sup_ids || !qry_ids ||
        !sup_y || !qry_y || !head_class || !sup_rows || !qry_rows) {
        fumi_set_error("fumi_sampler_next: null/empty argument");
        return FUMI_ERR_ARG;
    }
    const int N = s->N, K = s->K, Q = s->Q;
    // ---- stream 1: all B class tuples first (BatchSampler pulls the indices before any fetch)
    MT19937 py;
    std::memcpy(py.key, py_state, sizeof(py.key));
    py.pos = int(py_state[624]);
    py.retemper();
    for (int64_t b = 0; b < B; ++b) py_sample_range(py, s->C, N, classes + b * N);
    std::memcpy(py_state, py.key, sizeof(py.key));
    py_state[624] = uint32_t(py.pos);
    // class-size check (ClassSplitter_ raises ValueError)
    for (int64_t i = 0; i < B * N; ++i) {
        int64_t c = classes[i];
        int64_t n_c = s->offsets[c + 1] - s->offsets[c];
        if (n_c < K + Q) {
            fumi_set_error("The number of samples for one class (" + std::to_string(n_c) +
                           ") is smaller than the minimum number of samples per class required (" +
                           std::to_string(K + Q) + ").");
            return FUMI_ERR_DATA;
        }
    }
    // ---- stream 2a: hash-seeded permutations, independent per (task, class): thread pool
    std::vector<int64_t> head(size_t(B) * N * (K + Q));  // first K+Q entries of each permutation
    int nt = num_threads > 0 ? num_threads : int(std::thread::hardware_concurrency());
    nt = std::max(1, std::min<int>(nt, int(std::min<int64_t>(B, 64))));
    auto work = [&](int64_t b0, int64_t b1) {
        std::vector<int64_t> perm;
        MT19937 g;
        for (int64_t b = b0; b < b1; ++b) {
            const int64_t h = fumi_py_tuple_hash(classes + b * N, N);
            for (int p = 0; p < N; ++p) {
                const int64_t c = classes[b * N + p];
                const int64_t n_c = s->offsets[c + 1] - s->offsets[c];
                const uint32_t seed = uint32_t(uint64_t(h) + uint64_t(c));   // (hash + c + 0) % 2**32
                g.seed(seed);
                perm.resize(n_c);
                for (int64_t i = 0; i < n_c; ++i) perm[i] = i;
                np_shuffle(g, perm.data(), n_c);
                std::memcpy(&head[(size_t(b) * N + p) * (K + Q)], perm.data(), sizeof(int64_t) * (K + Q));
            }
        }
    };
    if (nt == 1) {
        work(0, B);
    } else {
        std::vector<std::thread> pool;
        int64_t per = (B + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) {
            int64_t b0 = t * per, b1 = std::min<int64_t>(B, b0 + per);
            if (b0 < b1) pool.emplace_back(work, b0, b1);
        }
        for (auto& th : pool) th.join();
    }
    // ---- stream 2b (shared RandomState(0), sequential) + stream 3 (torch randperm per task)
    TorchMT tg{torch_state};
    for (int64_t b = 0; b < B; ++b) {
        for (int p = 0; p < N; ++p) {
            const int64_t c = classes[b * N + p];
            const int64_t* ids_c = s->ids.data() + s->offsets[c];
            int64_t* hp = &head[(size_t(b) * N + p) * (K + Q)];
            np_shuffle(s->shared, hp, K);          // support picks, then query picks
            np_shuffle(s->shared, hp + K, Q);
            for (int k = 0; k < K; ++k) {
                sup_ids[(b * N + p) * K + k] = ids_c[hp[k]];
                sup_rows[(b * N + p) * K + k] = s->offsets[c] + hp[k];
            }
            for (int q = 0; q < Q; ++q) {
                qry_ids[(b * N + p) * Q + q] = ids_c[hp[K + q]];
                qry_rows[(b * N + p) * Q + q] = s->offsets[c] + hp[K + q];
            }
        }
    }
    for (int64_t b = 0; b < B; ++b) {
        int64_t* lp = label_perm + b * N;           // torch.randperm(N), CPU small-n path
        for (int i = 0; i < N; ++i) lp[i] = i;
        for (int i = 0; i < N - 1; ++i) {
            uint32_t z = tg.next() % uint32_t(N - i);
            std::swap(lp[i], lp[i + z]);
        }
        for (int p = 0; p < N; ++p) {
            for (int k = 0; k < K; ++k) sup_y[(b * N + p) * K + k] = lp[p];
            for (int q = 0; q < Q; ++q) qry_y[(b * N + p) * Q + q] = lp[p];
            head_class[b * N + lp[p]] = classes[b * N + p];
        }
    }
    return FUMI_OK;
}

// Sequential streams only; the per (task, class) permutations run on the device (sampler_expand.cu).
extern "C" int fumi_sampler_plan(fumi_sampler* s, int64_t B, uint32_t* py_state, uint32_t* torch_state,
                                 int64_t* classes, int64_t* label_perm, int64_t* head_class,
                                 uint32_t* perm_seed, int32_t* picks, int32_t* job_order) {
    if (!s || B <= 0 || !py_state || !torch_state || !classes || !label_perm || !head_class || !perm_seed || !picks) {
        fumi_set_error("fumi_sampler_plan: null/empty argument");
        return FUMI_ERR_ARG;
    }
    const int N = s->N, K = s->K, Q = s->Q;
    MT19937 py;                                        // stream 1: all B class tuples first
    std::memcpy(py.key, py_state, sizeof(py.key));
    py.pos = int(py_state[624]);
    py.retemper();
    for (int64_t b = 0; b < B; ++b) py_sample_range(py, s->C, N, classes + b * N);
    std::memcpy(py_state, py.key, sizeof(py.key));
    py_state[624] = uint32_t(py.pos);
    for (int64_t i = 0; i < B * N; ++i) {
        const int64_t c = classes[i];
        const int64_t n_c = s->offsets[c + 1] - s->offsets[c];
        if (n_c < K + Q) {
            fumi_set_error("The number of samples for one class (" + std::to_string(n_c) +
                           ") is smaller than the minimum number of samples per class required (" +
                           std::to_string(K + Q) + ").");
            return FUMI_ERR_DATA;
        }
    }
    std::vector<int64_t> slot(size_t(std::max(K, Q)));
    for (int64_t b = 0; b < B; ++b) {
        const int64_t h = fumi_py_tuple_hash(classes + b * N, N);
        for (int p = 0; p < N; ++p) {
            perm_seed[b * N + p] = uint32_t(uint64_t(h) + uint64_t(classes[b * N + p]));
            int32_t* pk = picks + (b * N + p) * (K + Q);
            // stream 2b: the shared RandomState(0) shuffles move positions, whatever they hold
            for (int k = 0; k < K; ++k) slot[k] = k;
            np_shuffle(s->shared, slot.data(), K);
            for (int k = 0; k < K; ++k) pk[k] = int32_t(slot[k]);
            for (int q = 0; q < Q; ++q) slot[q] = K + q;
            np_shuffle(s->shared, slot.data(), Q);
            for (int q = 0; q < Q; ++q) pk[K + q] = int32_t(slot[q]);
        }
    }
    TorchMT tg{torch_state};                           // stream 3: torch.randperm(N) per task
    for (int64_t b = 0; b < B; ++b) {
        int64_t* lp = label_perm + b * N;
        for (int i = 0; i < N; ++i) lp[i] = i;
        for (int i = 0; i < N - 1; ++i) {
            uint32_t z = tg.next() % uint32_t(N - i);
            std::swap(lp[i], lp[i + z]);
        }
        for (int p = 0; p < N; ++p) head_class[b * N + lp[p]] = classes[b * N + p];
    }
    if (job_order) {
        // device jobs (task, class) longest class first: a job's cost is its class size, and the expand kernel's time is
        // otherwise set by a long job that happens to start last (counting sort, stable)
        int64_t max_n = 0;
        for (int64_t c = 0; c < s->C; ++c) max_n = std::max(max_n, s->offsets[c + 1] - s->offsets[c]);
        std::vector<int32_t> start(size_t(max_n) + 2, 0);
        for (int64_t j = 0; j < B * N; ++j) ++start[size_t(max_n - (s->offsets[classes[j] + 1] - s->offsets[classes[j]])) + 1];
        for (size_t k = 1; k < start.size(); ++k) start[k] += start[k - 1];
        for (int64_t j = 0; j < B * N; ++j)
            job_order[start[size_t(max_n - (s->offsets[classes[j] + 1] - s->offsets[classes[j]]))]++] = int32_t(j);
    }
    return FUMI_OK;
}

// CPU emulator of k_fused_lanes (test infrastructure): runs rodio_b200/csrc/rb_lanes_core.h -- the very source the
// device kernel is compiled from -- on 32 host threads per warp (see rb_simt.h, RB_SIMT_EMULATE) and adds the
// per-warp partial rows in warp order like k_sum_groups.  Built and loaded by tests/test_lanes_emulator.py.
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

#include "warp_variants.h"
#include "../../rodio_b200/csrc/rb_lanes_plan.h"

namespace {
// the class decides the instantiation -- source channels ci, mixer channels co, PASS when from == to -- like the device launcher
void run_group_any(uint32_t ci, uint32_t co, const lanes::Args& a, uint32_t g, simt::WarpEmu* w, float* ring, bool hasb, bool ff2, bool npost,
                   bool pre = false, bool front = false, bool guard = false) {
    emu_run_group(ci, co, a, g, w, ring, hasb, ff2, npost, pre, front, guard);
}
constexpr int MAX_RS = lanes::Geo<2, 4>::RS;   // the largest ring of any variant
}  // namespace

// coverage of the last runs: [0] fast tiles, [1] slow tiles, [2] ring refills; reset = 1 clears the counters
extern "C" void rb_lanes_emu_counters(uint64_t* out, int reset) {
    for (int i = 0; i < 4; i++) out[i] = simt::g_emu_count[i];
    if (reset)
        for (int i = 0; i < 4; i++) simt::g_emu_count[i] = 0;
}

extern "C" int rb_lanes_emulate(const float* const* pcm, const uint64_t* n_frames, const uint64_t* out_len,
                                const uint64_t* mix_start, const float* coefs /* [n][5] b0 b1 b2 a1 a2 */,
                                const float* post, uint32_t n_rows, uint32_t channels /* mixer */, const uint32_t* ch_in /* per stream */,
                                const uint32_t* from, const uint32_t* to, uint64_t mix_len, int hasb, int want_ff2, int npost, float* out_mix, float* out_partials /* may be NULL */,
                                int* used_ff2, uint32_t* n_unsafe, const float* pre /* NULL: no gain in front of the conversion */,
                                int front /* the filter sits in front of the conversion */, const float* mid /* front: gain behind the filter, may be NULL */) {
    using namespace lanes;
    if (n_rows == 0 || (channels != 1 && channels != 2)) return 1;
    for (uint32_t r = 0; r < n_rows; r++)
        if (from[r] == 0 || from[r] > (1u << 20) || to[r] > (1u << 20) || !(ch_in[r] == channels || (ch_in[r] == 1 && channels == 2))) return 1;
    const uint32_t C = channels;   // n_frames / out_len / mix_start / mix_len count FRAMES; out_mix holds frames * C floats,
                                   // pcm[r] frames * ch_in[r] floats
    const float nan = std::numeric_limits<float>::quiet_NaN();
    simt::WarpEmu warp;
    // inputs: 16-byte aligned copies with a 16-byte tail pad of NaN (reading the pad as data would show)
    std::vector<std::vector<float>> store(n_rows);
    std::vector<Row> all(n_rows);
    bool ff2 = want_ff2 && hasb && !front;
    *n_unsafe = 0;
    for (uint32_t r = 0; r < n_rows; r++) {
        const uint32_t ci = ch_in[r];
        store[r].assign(n_frames[r] * ci + 4 + 4, nan);
        float* base = store[r].data();
        while ((uintptr_t)base & 15) base++;
        if (n_frames[r]) std::memcpy(base, pcm[r], n_frames[r] * ci * 4);
        warp.readable.push_back({(const char*)base, (const char*)(base + n_frames[r] * ci + 4)});
        Row& row = all[r];
        std::memset(&row, 0, sizeof(row));
        row.in = base, row.L = n_frames[r], row.out_len = out_len[r], row.mix_start = mix_start[r];
        row.n_int = n_interp(row.L, from[r], to[r], row.out_len);
        if (hasb) {
            const float* c = coefs + 5 * r;
            row.b0 = c[0], row.b1 = c[1], row.b2 = c[2], row.a1 = c[3], row.a2 = c[4];
            float k = 0;
            if (ff2_coeffs(row.b0, row.b1, row.b2, &k)) row.ffk = k;
            else ff2 = false;
        }
        row.post = npost ? post[r] : 1.0f;
        row.pre = pre ? pre[r] : 1.0f;
        row.mid = mid ? mid[r] : 1.0f;
        if (pre && !front && !pre_gain_keeps_class(row.pre)) row.flags |= ROW_FORCE_SLOW;
        bool ok = true;
        for (uint64_t i = 0; i < n_frames[r] * ci && ok; i++) ok = sample_in_class(pcm[r][i]);
        if (!ok) row.flags |= ROW_UNSAFE, (*n_unsafe)++;
    }
    *used_ff2 = ff2;
    alignas(16) static float zeros[CHUNK * 2] = {0};
    warp.readable.push_back({(const char*)zeros, (const char*)(zeros + CHUNK * C)});
    // one launch per rate pair over its own rows and partial rows
    const auto classes = classes_by_ratio(from, to, ch_in, n_rows);
    std::vector<Row> rows;
    std::vector<Args> launches;
    std::vector<uint32_t> launch_ci;
    std::vector<uint8_t> launch_guard;   // the class holds a row whose gain in front is out of the unguarded tile's range
    uint32_t n_groups_total = 0;
    const uint64_t pstride = round_up_tile(mix_len * C);
    for (const auto& cls : classes) {
        Args a{};
        a.n_rows = (uint32_t)cls.size(), a.n_groups = (a.n_rows + 31) / 32;
        fill_ratio(a, from[cls[0]], to[cls[0]], C);
        a.mix_len = mix_len, a.pstride = pstride;
        a.rows = (const Row*)(uintptr_t)rows.size();        // offsets until the vectors stop growing
        a.partial = (float*)(uintptr_t)n_groups_total;
        for (uint32_t i : cls) rows.push_back(all[i]);
        n_groups_total += a.n_groups;
        launches.push_back(a);
        launch_ci.push_back(ch_in[cls[0]]);
        bool guard = false;
        for (uint32_t i : cls) guard = guard || (all[i].flags & ROW_FORCE_SLOW);
        launch_guard.push_back(guard);
    }
    std::vector<float> partial((size_t)n_groups_total * pstride, 0.0f);
    std::vector<float> ring_store(32 * MAX_RS + 4, nan);
    float* ring = ring_store.data();
    while ((uintptr_t)ring & 15) ring++;
    for (size_t k = 0; k < launches.size(); k++) {
        Args& a = launches[k];
        a.rows = rows.data() + (uintptr_t)a.rows, a.partial = partial.data() + (uintptr_t)a.partial * pstride, a.zeros = zeros;
        for (uint32_t g = 0; g < a.n_groups; g++) {
            for (int i = 0; i < 32 * MAX_RS; i++) ring[i] = nan;
            run_group_any(launch_ci[k], channels, a, g, &warp, ring, hasb, ff2, npost, pre != nullptr, front != 0, launch_guard[k] != 0);
        }
    }
    for (uint64_t m = 0; m < mix_len * C; m++) {
        float acc = 0.0f;
        for (uint32_t g = 0; g < n_groups_total; g++) acc = acc + partial[(size_t)g * pstride + m];
        out_mix[m] = acc;
    }
    if (out_partials) std::memcpy(out_partials, partial.data(), partial.size() * 4);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Streaming session on the emulator: the bookkeeping of rodio_b200/csrc/rb_session_plan.h driving the same warp
// program block by block.  ops: triples (kind, stream, count) -- kind 0: push `count` more frames of the stream's PCM
// (the stream ends when all of it has been pushed, or on kind 2), kind 1: render up to `count` mixer frames,
// kind 2: mark the stream ended now (whatever was pushed is all there is), kind 3: the stream's gain becomes the float
// whose bits are `count` from the next block on (rb_session_set_amplify), kind 4: Mixer::add of a source that was declared
// held (mix_start == ~0 in the arguments): it joins at the frame rendered next.  After the last op everything is ended and
// drained (sources still held are dropped).  Returns the number of mixer frames written to out (capacity out_cap), or -1.
#include "../../rodio_b200/csrc/rb_session_plan.h"

extern "C" long long rb_session_emulate(const float* const* pcm, const uint64_t* n_frames, const uint64_t* mix_start,
                                        const float* coefs, const float* post, uint32_t n_rows, uint32_t channels /* mixer */,
                                        const uint32_t* ch_in, const uint32_t* from, const uint32_t* to, int hasb, int npost, const uint64_t* ops,
                                        uint64_t n_ops, float* out, uint64_t out_cap, uint64_t* n_renders,
                                        uint64_t* pushed_total /* [n_rows] */, uint64_t* joined_at /* [n_rows]: mixer frame, ~0 = never */) {
    using namespace lanes;
    if (n_rows == 0 || (channels != 1 && channels != 2)) return -1;
    for (uint32_t r = 0; r < n_rows; r++)
        if (from[r] == 0 || from[r] > (1u << 20) || to[r] > (1u << 20) || !(ch_in[r] == channels || (ch_in[r] == 1 && channels == 2))) return -1;
    const auto classes = classes_by_ratio(from, to, ch_in, n_rows);   // fixed for the session: rows are laid out class by class
    const uint32_t C = channels;   // frames everywhere; out holds frames * C floats, pcm[r] / FIFO r frames * ch_in[r] floats
    const float nan = std::numeric_limits<float>::quiet_NaN();
    simt::WarpEmu warp;
    std::vector<session::Stream> st(n_rows);
    std::vector<std::vector<float>> fifo_store(n_rows);
    std::vector<float*> fifo(n_rows);
    std::vector<uint8_t> unsafe(n_rows, 0);
    std::vector<float> state_store(4 * C * n_rows + 4, 0.0f);
    float* state = state_store.data();
    while ((uintptr_t)state & 15) state++;
    bool ff2 = hasb;
    std::vector<float> ffk(n_rows, 0.0f);
    std::vector<float> gain(n_rows, 1.0f);
    for (uint32_t r = 0; r < n_rows; r++) {
        if (npost) gain[r] = post[r];
        fifo_store[r].assign(n_frames[r] * ch_in[r] + 16, nan);
        fifo[r] = fifo_store[r].data();
        while ((uintptr_t)fifo[r] & 15) fifo[r]++;
        warp.readable.push_back({(const char*)fifo[r], (const char*)(fifo[r] + n_frames[r] * ch_in[r] + 8)});
        st[r].mix_start = mix_start[r], st[r].from = from[r], st[r].to = to[r];
        if (mix_start[r] == ~0ull) st[r].held = true, st[r].mix_start = 0;
        if (hasb) {
            const float* c = coefs + 5 * r;
            if (!ff2_coeffs(c[0], c[1], c[2], &ffk[r])) ff2 = false;
        }
    }
    alignas(16) static float zeros[CHUNK * 2] = {0};
    warp.readable.push_back({(const char*)zeros, (const char*)(zeros + CHUNK * C)});
    std::vector<float> ring_store(32 * MAX_RS + 4, nan);
    float* ring = ring_store.data();
    while ((uintptr_t)ring & 15) ring++;
    uint64_t T = 0, written = 0;
    *n_renders = 0;

    auto push = [&](uint32_t r, uint64_t n) {
        session::Stream& s = st[r];
        n = std::min(n, n_frames[r] - s.pushed);
        const uint32_t ci = ch_in[r];
        for (uint64_t k = 0; k < n * ci; k++) {
            const float v = pcm[r][s.pushed * ci + k];
            fifo[r][s.fill() * ci + k] = v;
            if (!sample_in_class(v)) unsafe[r] = 1;
        }
        s.pushed += n;
        if (s.pushed == n_frames[r]) s.eof = true;
    };
    auto render = [&](uint64_t max_frames) -> bool {   // false: session ended
        session::resolve_queue(st, T);
        bool ended = false;
        const uint64_t n = session::renderable(st, T, max_frames, &ended);
        if (ended) return false;
        if (n == 0) return true;
        if ((written + n) * C > out_cap) std::abort();
        std::vector<Row> all(n_rows);
        std::vector<session::Part> parts(n_rows);
        for (uint32_t r = 0; r < n_rows; r++) {
            parts[r] = session::part_of(st[r], T, n);
            Row& row = all[r];
            std::memset(&row, 0, sizeof(row));
            row.in = fifo[r], row.L = st[r].fill(), row.out_len = parts[r].out_len, row.mix_start = parts[r].mix_start;
            row.n_int = parts[r].n_int, row.o0 = parts[r].o0, row.i0 = st[r].i0, row.state = state + 4 * C * r;
            if (hasb) {
                const float* c = coefs + 5 * r;
                row.b0 = c[0], row.b1 = c[1], row.b2 = c[2], row.a1 = c[3], row.a2 = c[4], row.ffk = ffk[r];
            }
            row.post = gain[r], row.pre = 1.0f, row.mid = 1.0f;   // no gain in front (the DOWN variants always apply Row::pre)
            row.flags = (unsafe[r] ? ROW_UNSAFE : 0u) | (parts[r].continues ? ROW_CONTINUES : 0u);
        }
        const uint64_t pstride = round_up_tile(n * C);
        uint32_t n_groups_total = 0;
        for (const auto& cls : classes) n_groups_total += ((uint32_t)cls.size() + 31) / 32;
        std::vector<float> partial((size_t)n_groups_total * pstride, 0.0f);
        std::vector<Row> rows;
        rows.reserve(n_rows);
        uint32_t g0 = 0;
        for (const auto& cls : classes) {           // one launch per rate pair
            Args a{};
            a.n_rows = (uint32_t)cls.size(), a.n_groups = (a.n_rows + 31) / 32;
            fill_ratio(a, from[cls[0]], to[cls[0]], C);
            a.mix_len = n, a.pstride = pstride;
            const size_t first = rows.size();
            for (uint32_t i : cls) rows.push_back(all[i]);
            a.rows = rows.data() + first, a.partial = partial.data() + (size_t)g0 * pstride, a.zeros = zeros;
            for (uint32_t g = 0; g < a.n_groups; g++) {
                for (int i = 0; i < 32 * MAX_RS; i++) ring[i] = nan;
                run_group_any(ch_in[cls[0]], channels, a, g, &warp, ring, hasb, ff2, npost);
            }
            g0 += a.n_groups;
        }
        for (uint64_t m = 0; m < n * C; m++) {
            float acc = 0.0f;
            for (uint32_t g = 0; g < n_groups_total; g++) acc = acc + partial[(size_t)g * pstride + m];
            out[written * C + m] = acc;
        }
        written += n, T += n, (*n_renders)++;
        for (uint32_t r = 0; r < n_rows; r++) {
            const uint64_t fill_before = st[r].fill();
            const uint64_t drop = session::advance(st[r], parts[r]);
            if (drop) {
                const uint32_t ci = ch_in[r];
                std::memmove(fifo[r], fifo[r] + drop * ci, (fill_before - drop) * ci * sizeof(float));
                for (uint64_t k = (fill_before - drop) * ci; k < fill_before * ci; k++) fifo[r][k] = nan;
            }
        }
        return true;
    };

    for (uint64_t k = 0; k < n_ops; k++) {
        const uint64_t kind = ops[3 * k], r = ops[3 * k + 1], cnt = ops[3 * k + 2];
        if (kind == 0) push((uint32_t)r, cnt);
        else if (kind == 1) render(cnt);
        else if (kind == 2) st[r].eof = true;
        else if (kind == 4) session::start(st[r], T);
        else if (kind == 5) st[r].follows = (int64_t)cnt;
        else {
            const uint32_t bits = (uint32_t)cnt;
            std::memcpy(&gain[r], &bits, 4);
        }
    }
    for (uint32_t r = 0; r < n_rows; r++) st[r].eof = true;   // whatever was pushed is all there is
    session::resolve_queue(st, T);
    for (uint32_t r = 0; r < n_rows; r++) pushed_total[r] = st[r].pushed;
    for (uint32_t r = 0; r < n_rows; r++) {
        joined_at[r] = st[r].held ? ~0ull : st[r].mix_start;
        if (st[r].held) st[r].held = false, st[r].pushed = st[r].i0 = 0, st[r].out_done = 0;   // never added: contributes nothing
    }
    while (render(1ull << 20)) {
        bool ended = false;
        if (session::renderable(st, T, 1, &ended) == 0 && !ended) std::abort();   // no progress
    }
    return (long long)written;
}

// ------------------------------------------------------------------------------------------------------------------
// Plan-only fuzz of rb_session_plan.h (no kernel): random pushes / renders / early ends on random rate pairs; checks
// the invariants the device code relies on.  Returns 0, or a line number of the violated check.
extern "C" int rb_session_plan_fuzz(uint64_t seed, uint32_t n_cases) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&](uint64_t n) { x ^= x << 13, x ^= x >> 7, x ^= x << 17; return n ? x % n : 0; };
#define CHECK(c) do { if (!(c)) return __LINE__; } while (0)
    for (uint32_t cs = 0; cs < n_cases; cs++) {
        const uint32_t ns = 1 + (uint32_t)rnd(5);
        const uint64_t cap = 64 + rnd(3000);
        std::vector<session::Stream> st(ns);
        std::vector<uint64_t> total(ns), rendered(ns, 0);
        for (uint32_t r = 0; r < ns; r++) {
            uint32_t to = 1 + (uint32_t)rnd(400), from = rnd(4) ? 1 + (uint32_t)rnd(to) : 1 + (uint32_t)rnd(3 * to);   // mostly from <= to (1:1 included), some down-sampling
            const uint32_t g = std::gcd(from, to);
            st[r].from = from / g, st[r].to = to / g;
            st[r].mix_start = rnd(3) ? 0 : rnd(500), total[r] = rnd(6000);
            if (rnd(4) == 0) st[r].held = true, st[r].mix_start = 0;     // Mixer::add comes later
        }
        uint64_t T = 0;
        for (int step = 0; step < 4000; step++) {
            const uint32_t r = (uint32_t)rnd(ns);
            if (st[r].held && rnd(12) == 0) session::start(st[r], T);
            if (!st[r].eof) {
                uint64_t n = std::min<uint64_t>(rnd(900), total[r] - st[r].pushed);
                n = std::min<uint64_t>(n, cap - st[r].fill());            // the device refuses more than the FIFO takes
                st[r].pushed += n;
                if (st[r].pushed == total[r] || rnd(400) == 0) st[r].eof = true;
            }
            if (rnd(3) == 0) continue;
            bool ended = false;
            const uint64_t n = session::renderable(st, T, 1 + rnd(1500), &ended);
            if (ended) break;
            for (uint32_t q = 0; q < ns; q++) {
                const uint32_t from = st[q].from, to = st[q].to;
                const session::Part p = session::part_of(st[q], T, n);
                CHECK(p.mix_start + p.out_len <= n);
                if (p.out_len) {
                    CHECK(p.o0 == st[q].out_done);                                   // no output skipped or repeated
                    CHECK(p.mix_start + T == st[q].mix_start + p.o0);                // it sits where the timeline says
                    const uint64_t first = (p.o0 * (uint64_t)from) / to, last = ((p.o0 + p.out_len - 1) * (uint64_t)from) / to;
                    CHECK(first >= st[q].i0 && last < st[q].pushed);                  // taps inside the FIFO ...
                    CHECK(p.n_int <= p.out_len);
                    if (p.n_int) CHECK(((p.o0 + p.n_int - 1) * (uint64_t)from) / to + 1 < st[q].pushed);   // ... right taps too
                    if (p.n_int < p.out_len) CHECK(st[q].eof && p.out_len - p.n_int == 1);                  // only the raw last frame
                    if (!session::finished(st[q]) && st[q].mix_start + st[q].out_done <= T) CHECK(p.mix_start == 0);
                } else {
                    CHECK(session::finished(st[q]) || st[q].held || st[q].mix_start >= T + n || n == 0);
                }
                const uint64_t fill = st[q].fill();
                const uint64_t drop = session::advance(st[q], p);
                rendered[q] += p.out_len;
                CHECK(drop <= fill && (st[q].i0 & 3) == 0);
                CHECK(st[q].eof || (st[q].out_done * (uint64_t)from) / to >= st[q].i0);
                // what stays in the FIFO is bounded: an unfinished stream keeps at most the frames not yet usable + 4
                CHECK(p.continues == !session::finished(st[q]) || p.out_len == 0);
                if (st[q].held) CHECK(p.out_len == 0 && drop == 0 && st[q].out_done == 0);
            }
            T += n;
        }
        for (uint32_t r = 0; r < ns; r++) st[r].eof = true;
        for (int guard = 0; guard < 100000; guard++) {
            for (uint32_t r = 0; r < ns; r++)
                if (st[r].held && rnd(3) == 0) session::start(st[r], T);        // the held ones join as the others drain
            bool ended = false;
            const uint64_t n = session::renderable(st, T, 1 + rnd(5000), &ended);
            if (ended) break;
            bool waiting = false;
            for (uint32_t r = 0; r < ns; r++) waiting = waiting || st[r].held;
            CHECK(n > 0 || waiting);                                             // only a held source may stall the drain
            for (uint32_t q = 0; q < ns; q++) {
                const session::Part p = session::part_of(st[q], T, n);
                rendered[q] += p.out_len;
                session::advance(st[q], p);
            }
            T += n;
        }
        for (uint32_t r = 0; r < ns; r++) CHECK(rendered[r] == session::out_total(st[r].pushed, st[r].from, st[r].to));
    }
#undef CHECK
    return 0;
}

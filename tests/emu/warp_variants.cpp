// See warp_variants.h.  RB_EMU_PART = 0..11: bit 0 PRE, bit 1 PASS, bits 2.. the channel layout (0: 1->1, 1: 1->2, 2: 2->2);
// RB_EMU_PART = 12: the dispatcher.
#include "warp_variants.h"
#include "../../rodio_b200/csrc/rb_lanes_plan.h"

#ifndef RB_EMU_PART
#error "compile with -DRB_EMU_PART=0..12"
#endif

#define RB_EMU_DECL(k) void emu_run_part_##k(const lanes::Args&, uint32_t, simt::WarpEmu*, float*, bool, bool, bool, bool, bool)

#if RB_EMU_PART < 12
namespace {
constexpr bool PRE = (RB_EMU_PART & 1) != 0, PASS = (RB_EMU_PART & 2) != 0;
constexpr int LAYOUT = RB_EMU_PART >> 2, CI = LAYOUT == 2 ? 2 : 1, CO = LAYOUT == 0 ? 1 : 2;
template <bool HASB, bool FF2, int NPOST, bool FRONT = false, bool DOWN = false, bool GUARD = false>
void run(const lanes::Args& a, uint32_t group, simt::WarpEmu* w, float* ring) {
    simt::run_warp(w, [&] { lanes::warp_main<CI, CO, HASB, FF2, NPOST, PASS, PRE, FRONT, DOWN, GUARD>(a, group, ring); });
}
}  // namespace
#define RB_EMU_CAT2(a, b) a##b
#define RB_EMU_CAT(a, b) RB_EMU_CAT2(a, b)
void RB_EMU_CAT(emu_run_part_, RB_EMU_PART)(const lanes::Args& a, uint32_t g, simt::WarpEmu* w, float* ring, bool hasb, bool ff2, bool npost, bool front, bool guard) {
    if constexpr (PRE && !PASS) {   // a gain in front out of the unguarded tile's range: every quotient is checked
        if (guard) {
            if (hasb && ff2 && npost) run<true, true, 1, false, false, true>(a, g, w, ring);
            else if (hasb && ff2) run<true, true, 0, false, false, true>(a, g, w, ring);
            else if (hasb && npost) run<true, false, 1, false, false, true>(a, g, w, ring);
            else if (hasb) run<true, false, 0, false, false, true>(a, g, w, ring);
            else if (npost) run<false, false, 1, false, false, true>(a, g, w, ring);
            else run<false, false, 0, false, false, true>(a, g, w, ring);
            return;
        }
    }
    if constexpr (!PRE) {   // the filter in front of the conversion: plain coefficients, the gain in front always applied
        if (front) {
            if constexpr (!PASS) {
                if (lanes::ratio_runs_down(a.from, a.to)) {
                    npost ? run<true, false, 1, true, true>(a, g, w, ring) : run<true, false, 0, true, true>(a, g, w, ring);
                    return;
                }
            }
            npost ? run<true, false, 1, true>(a, g, w, ring) : run<true, false, 0, true>(a, g, w, ring);
            return;
        }
    }
    if constexpr (!PRE && !PASS) {   // sources above the mixer's rate, up to twice: fast tiles of their own (rb_lanes_plan.h)
        if (lanes::ratio_runs_down(a.from, a.to)) {
            if (hasb && ff2 && npost) run<true, true, 1, false, true>(a, g, w, ring);
            else if (hasb && ff2) run<true, true, 0, false, true>(a, g, w, ring);
            else if (hasb && npost) run<true, false, 1, false, true>(a, g, w, ring);
            else if (hasb) run<true, false, 0, false, true>(a, g, w, ring);
            else if (npost) run<false, false, 1, false, true>(a, g, w, ring);
            else run<false, false, 0, false, true>(a, g, w, ring);
            return;
        }
    }
    if (hasb && ff2 && npost) run<true, true, 1>(a, g, w, ring);
    else if (hasb && ff2) run<true, true, 0>(a, g, w, ring);
    else if (hasb && npost) run<true, false, 1>(a, g, w, ring);
    else if (hasb) run<true, false, 0>(a, g, w, ring);
    else if (npost) run<false, false, 1>(a, g, w, ring);
    else run<false, false, 0>(a, g, w, ring);
}
#else
RB_EMU_DECL(0); RB_EMU_DECL(1); RB_EMU_DECL(2); RB_EMU_DECL(3); RB_EMU_DECL(4); RB_EMU_DECL(5);
RB_EMU_DECL(6); RB_EMU_DECL(7); RB_EMU_DECL(8); RB_EMU_DECL(9); RB_EMU_DECL(10); RB_EMU_DECL(11);
void emu_run_group(uint32_t ci, uint32_t co, const lanes::Args& a, uint32_t group, simt::WarpEmu* w, float* ring, bool hasb, bool ff2,
                   bool npost, bool pre, bool front, bool guard) {
    using Fn = void (*)(const lanes::Args&, uint32_t, simt::WarpEmu*, float*, bool, bool, bool, bool, bool);
    static const Fn parts[12] = {emu_run_part_0, emu_run_part_1, emu_run_part_2, emu_run_part_3, emu_run_part_4,  emu_run_part_5,
                                 emu_run_part_6, emu_run_part_7, emu_run_part_8, emu_run_part_9, emu_run_part_10, emu_run_part_11};
    const uint32_t layout = ci == 2 ? 2u : (co == 2 ? 1u : 0u);
    const bool down = !front && lanes::ratio_runs_down(a.from, a.to);   // the DOWN variants apply Row::pre themselves
    parts[layout * 4 + (a.from == a.to ? 2u : 0u) + (pre && !front && !down ? 1u : 0u)](a, group, w, ring, hasb, ff2, npost, front, guard);
}
#endif

// The instantiations of lanes::warp_main (rodio_b200/csrc/rb_lanes_core.h) on the SIMT emulator, shared by the two harnesses
// (lanes_emu.cpp, hostemu.cpp).  They are compiled in twelve parts -- warp_variants.cpp with -DRB_EMU_PART=0..11, one per
// (source channels, mixer channels, PASS, PRE) -- so that the objects build in parallel (tests/emu/build_emu.py).
#pragma once
#define RB_SIMT_EMULATE 1
#include "../../rodio_b200/csrc/rb_lanes_core.h"

// Runs warp `group` of launch `a` like the device launcher picks its kernel: source channels ci, mixer channels co, PASS when
// from == to, PRE when the class carries a gain in front of the conversion, FRONT when its filter sits there as well, GUARD (with PRE) when a gain in front is out of the
// unguarded tile's range.
void emu_run_group(uint32_t ci, uint32_t co, const lanes::Args& a, uint32_t group, simt::WarpEmu* w, float* ring, bool hasb, bool ff2,
                   bool npost, bool pre, bool front = false, bool guard = false);

"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header
declares, host-only helpers match the oracle bit for bit, closed-form lengths match the literal
iterator, argument validation mirrors the reference's panics.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from rodio_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "rodio_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    names = header_functions()
    assert len(names) >= 30
    L = C.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rodio_b200.h but not exported"
    assert set(names) == set(capi.SYMBOLS), set(names) ^ set(capi.SYMBOLS)
    assert rb.lib().rb_abi_version() == 1


def test_struct_layout_matches_header(built):
    assert C.sizeof(capi.rb_effect) == 80
    assert C.sizeof(capi.rb_stream_desc) == 40
    assert capi.rb_stream_desc.effects.offset == 24 and capi.rb_stream_desc.mix_start.offset == 32


def test_no_gpu_fails_loudly(built):
    h = C.c_void_p()
    st = rb.lib().rb_context_create(0, C.byref(h))
    if st == capi.RB_OK:
        rb.lib().rb_context_destroy(h)
        pytest.skip("a GPU is present")
    assert st == capi.RB_ERR_CUDA
    assert b"no CPU fallback" in rb.lib().rb_last_error()
    with pytest.raises(rb.RodioB200Error):
        rb.Context(0)


def test_status_strings(built):
    for s in range(0, 10):
        assert rb.lib().rb_status_string(s)


def _out_len(n, f, t, c):
    v = C.c_uint64()
    st = rb.lib().rb_sample_rate_out_len(n, f, t, c, C.byref(v))
    return st, v.value


def test_sample_rate_out_len_matches_literal_iterator(built):
    rng = np.random.default_rng(7)
    rates = [8000, 11025, 16000, 22050, 44100, 48000, 88200, 96000, 176400, 192000, 352800, 384000, 39690, 40000,
             1, 2, 3, 7, 1000, 7000, 2400, 12000]
    for _ in range(400):
        f, t = int(rng.choice(rates)), int(rng.choice(rates))
        c = int(rng.integers(1, 5))
        frames = int(rng.integers(0, 700))
        if frames * t / f > 100_000:      # keep the literal drain short
            frames = int(100_000 * f / t)
        x = rng.uniform(-1, 1, frames * c).astype(np.float32)
        st, n = _out_len(x.size, f, t, c)
        assert st == capi.RB_OK
        assert n == oracle.sample_rate_converter(x, f, t, c).size, (f, t, c, frames)
    for f, t, L, want in [(147, 160, 147, 160), (44100, 48000, 1000, 1089), (44100, 48000, 44100, 48000),
                          (1, 1, 5, 5), (44100, 48000, 1, 1), (44100, 48000, 2, 3), (44100, 48000, 0, 0)]:
        assert _out_len(L, f, t, 1) == (capi.RB_OK, want)


def test_out_len_errors(built):
    assert _out_len(10, 0, 48000, 1)[0] == capi.RB_ERR_INVALID_ARGUMENT
    assert _out_len(10, 48000, 48000, 0)[0] == capi.RB_ERR_INVALID_ARGUMENT
    assert _out_len(7, 44100, 48000, 2)[0] == capi.RB_ERR_UNALIGNED_FRAMES
    assert _out_len(10, 96001, 192000, 1)[0] == capi.RB_ERR_RATIO_OVERFLOW
    v = C.c_uint64()
    assert rb.lib().rb_channels_out_len(6, 3, 2, C.byref(v)) == capi.RB_OK and v.value == 4
    assert rb.lib().rb_channels_out_len(6, 0, 2, C.byref(v)) == capi.RB_ERR_INVALID_ARGUMENT
    assert rb.lib().rb_channels_out_len(7, 2, 1, C.byref(v)) == capi.RB_ERR_UNALIGNED_FRAMES


def test_host_helpers_bit_exact_with_oracle(built):
    L, O = rb.lib(), oracle.lib()
    rng = np.random.default_rng(11)
    for _ in range(300):
        rate = int(rng.integers(1, 400000))
        fac = float(np.float32(rng.uniform(0.0, 4.0)))
        assert L.rb_speed_sample_rate(rate, fac) == O.ro_speed_sample_rate(rate, fac)
        ns = int(rng.integers(0, 5_000_000_000))
        ch = int(rng.integers(1, 9))
        assert L.rb_delay_samples(ns, rate, ch) == O.ro_delay_samples(ns, rate, ch)
        db = float(np.float32(rng.uniform(-100, 100)))
        assert np.float32(L.rb_db_to_linear(db)).view(np.uint32) == np.float32(O.ro_db_to_linear(db)).view(np.uint32)
        lin = float(np.float32(rng.uniform(1e-6, 100)))
        assert np.float32(L.rb_linear_to_db(lin)).view(np.uint32) == np.float32(O.ro_linear_to_db(lin)).view(np.uint32)
    assert L.rb_speed_sample_rate(44100, 0.9) == 39690
    assert L.rb_speed_sample_rate(5, 0.0) == 1                  # .max(1.0)
    assert L.rb_delay_samples(50_000_000, 48000, 2) == 4800


def test_spatial_volumes_bit_exact_with_oracle(built):
    rng = np.random.default_rng(12)
    for _ in range(200):
        e, l, r = (rng.uniform(-5, 5, 3).astype(np.float32) for _ in range(3))
        out = np.zeros(2, np.float32)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        rb.lib().rb_spatial_volumes(fp(e), fp(l), fp(r), fp(out))
        assert np.array_equal(out.view(np.uint32), oracle.spatial_volumes(e, l, r).view(np.uint32))


def test_duration_from_secs_f32():
    assert rb.Duration.from_secs_f32(0.5) == 500_000_000
    assert rb.Duration.from_secs_f32(2.0) == 2_000_000_000
    # 0.05f32 = 0.0500000007450580596923828125 -> 50_000_001 ns (rounded to nearest)
    assert rb.Duration.from_secs_f32(0.05) == 50_000_001
    assert rb.Duration.from_millis(5) == 5_000_000


def test_source_metadata_follows_the_trait():
    s = rb.SamplesBuffer(2, 44100, np.zeros(8, np.float32))
    assert (s.channels(), s.sample_rate(), s.span_len) == (2, 44100, 8)
    assert s.speed(0.9).sample_rate() == 39690 and s.speed(0.9).channels() == 2
    sp = rb.Spatial(s, [0, 1, 0], [-1, 0, 0], [1, 0, 0])
    assert sp.channels() == 2 and sp.sample_rate() == 44100
    u = rb.UniformSourceIterator(s.amplify(2.0), 1, 48000)
    assert (u.channels(), u.sample_rate()) == (1, 48000)
    assert rb.TestSource(np.zeros(4, np.float32), 1, 48000).span_len == 0
    with pytest.raises(ValueError):
        rb.SamplesBuffer(0, 44100, [])
    with pytest.raises(ValueError):
        rb.mixer(1, 0)
    _, out = rb.mixer(2, 48000)
    assert out.channels() == 2 and out.sample_rate() == 48000 and out.current_span_len() is None
    with pytest.raises(rb.RodioB200Error):
        out.try_seek(0)


# ---- planner closed forms vs the literal pull iterator (host only) --------------------------
from chains import CHAINS, LIMIT_CHAINS  # noqa: E402
from helpers import noise, to_oracle  # noqa: E402


@pytest.mark.parametrize("name", sorted(CHAINS) + sorted(LIMIT_CHAINS))
def test_planner_lengths_match_literal_iterator(built, name):
    src = {**CHAINS, **LIMIT_CHAINS}[name]()
    want, ch, rate = oracle.chain(to_oracle(src))
    for mix in [(1, 48000), (2, 48000), (2, 44100), (3, 32000)]:
        n, pch, prate, cn = rb.plan(src, *mix)
        assert (pch, prate, cn) == (ch, rate, want.size), name
        assert n == oracle.chain_uniform(to_oracle(src), *mix).size, (name, mix)


def test_planner_random_uniform_lengths(built):
    """Spans, partial trailing frames, channel up/down-mix: closed form == literal iterator."""
    rng = np.random.default_rng(31)
    rates = [8000, 11025, 22050, 44100, 48000, 96000, 39690, 40000]
    for i in range(300):
        c = int(rng.integers(1, 5))
        n = int(rng.integers(0, 5000))
        kind = int(rng.integers(0, 3))
        x = np.zeros(n, np.float32)
        if kind == 0:
            n -= n % c
            src = rb.SamplesBuffer(c, int(rng.choice(rates)), x[:n])                      # spans
        elif kind == 1:
            n -= n % c
            src = rb.Source(x[:n], c, int(rng.choice(rates)), span_len=int(c * rng.integers(1, 600)))
        else:
            n -= n % c
            src = rb.TestSource(x[:n], c, int(rng.choice(rates))).delay(rb.Duration.from_nanos(int(rng.integers(0, 90000))))
        mix = (int(rng.integers(1, 5)), int(rng.choice(rates)))
        try:
            got = rb.plan(src, *mix)[0]
        except rb.RodioB200Error as e:
            assert e.status in (capi.RB_ERR_UNSUPPORTED,), e
            continue
        assert got == oracle.chain_uniform(to_oracle(src), *mix).size, (i, kind, c, n, mix)


def test_planner_cfg1_shape(built):
    """BASELINE cfg1 (benches/resampler.rs): 12 s of s16 stereo 44.1 kHz -> take_duration(10 s) ->
    UniformSourceIterator(2 ch, 48 kHz), at full size, against the literal iterators of the oracle.
    take.rs:65-69 truncates the per-sample duration to 11 337 ns: the take yields 882 067 samples plus one of
    frame padding (882 068) -- and the padding is never pulled through the UniformSourceIterator,
    whose span-limited Take stops at the 882 067 the source reports (take.rs:171-196, uniform.rs:50-68)."""
    x = (noise(2 * 44100 * 12, 31, 0.9) * 30000).astype(np.int16)
    src = rb.TestSource(x, 2, 44100).take_duration(rb.Duration.from_secs(10))
    chain, ch, rate = oracle.chain(to_oracle(src))
    assert (chain.size, ch, rate) == (882_068, 2, 44100)
    n, pch, prate, cn = rb.plan(src, 2, 48000)
    assert (pch, prate, cn) == (2, 44100, 882_068)
    want = oracle.chain_uniform(to_oracle(src), 2, 48000)
    assert n == want.size == 960_076
    for to_rate in (8000, 44100, 96000, 192000):
        assert rb.plan(src, 2, to_rate)[0] == oracle.chain_uniform(to_oracle(src), 2, to_rate).size


def test_planner_take_ramp_distortion_lengths(built):
    """take_duration padding is never pulled through a UniformSourceIterator / reverb (take.rs:180-196)."""
    rng = np.random.default_rng(9)
    rates = [8000, 11025, 22050, 44100, 48000, 96000, 39690]
    for i in range(400):
        c = int(rng.integers(1, 5))
        n = int(rng.integers(0, 600))
        n -= n % c
        x, r = np.zeros(n, np.float32), int(rng.choice(rates))
        base = rb.SamplesBuffer(c, r, x) if i % 2 else rb.TestSource(x, c, r)
        d = rb.Duration.from_nanos(int(rng.integers(0, 30_000_000)))
        k = i % 5
        if k == 0:
            src = base.take_duration(d)
        elif k == 1:
            src = base.take_duration(d, True).amplify(0.5).low_pass(300)
        elif k == 2:
            src = base.fade_in(rb.Duration.from_millis(3)).take_duration(d).reverb(rb.Duration.from_micros(int(rng.integers(0, 900))), 0.3)
        elif k == 3:
            src = base.delay(rb.Duration.from_micros(int(rng.integers(0, 900)))).take_duration(d).distortion(2.0, 0.5)
        else:
            src = rb.UniformSourceIterator(base.take_duration(d), int(rng.integers(1, 4)), int(rng.choice(rates))) \
                .fade_out(rb.Duration.from_millis(2))
        mix = (int(rng.integers(1, 5)), int(rng.choice(rates)))
        got = rb.plan(src, *mix)
        w, ch, rate = oracle.chain(to_oracle(src))
        assert (got[1], got[2], got[3]) == (ch, rate, w.size), (i, k)
        assert got[0] == oracle.chain_uniform(to_oracle(src), *mix).size, (i, k)
    with pytest.raises(ValueError):
        rb.TestSource(np.zeros(4, np.float32), 1, 48000).fade_in(0)
    with pytest.raises(rb.RodioB200Error):
        rb.plan(rb.TestSource(np.zeros(4, np.float32), 1, 48000).distortion(2.0, -1.0), 1, 48000)


def test_streams_plan_survives_random_descriptor_graphs(built):
    """rb_streams_plan on random descriptor arrays -- MIX / APPEND pointing anywhere (themselves, each other in cycles, out of
    range), consumed marks on and off, generators in odd places: every call returns a status, none crashes, and an accepted array
    never has a consumed descriptor that reaches the mixer."""
    import ctypes as C
    rng = np.random.default_rng(99)
    L = rb.lib()
    ok = 0
    for _ in range(3000):
        n = int(rng.integers(1, 6))
        descs = (capi.rb_stream_desc * n)()
        keep = []
        for i in range(n):
            k = int(rng.integers(0, 4))
            fx = (capi.rb_effect * max(1, k))()
            for j in range(k):
                kind = int(rng.choice([capi.RB_FX_AMPLIFY, capi.RB_FX_LOW_PASS, capi.RB_FX_MIX, capi.RB_FX_APPEND, capi.RB_FX_SIGNAL,
                                       capi.RB_FX_UNIFORM, capi.RB_FX_TAKE_DURATION, 99]))
                fx[j].kind = kind
                fx[j].u32[0] = int(rng.integers(0, n + 2))
                fx[j].u32[1] = int(rng.choice([0, 8000, 48000]))
                fx[j].f32[0] = float(rng.choice([0.5, 440.0, -1.0]))
                fx[j].ns[0] = int(rng.choice([0, 100, 10_000_000]))
            keep.append(fx)
            d = descs[i]
            d.sample_rate = int(rng.choice([0, 8000, 44100, 48000]))
            d.channels = int(rng.choice([0, 1, 2, 3]))
            d.format = capi.RB_FMT_F32
            d.n_samples = int(rng.choice([0, 1, 6, 600]))
            d.span_len = int(rng.choice([0, 0, d.n_samples, 7]))
            d.n_effects = k
            d.effects = C.cast(fx, C.POINTER(capi.rb_effect))
            d.mix_start = capi.RB_MIX_START_CONSUMED if rng.integers(0, 2) else int(rng.integers(0, 100))
        out_len = C.c_uint64()
        for i in range(n):
            st = L.rb_streams_plan(descs, n, i, 2, 48000, C.byref(out_len), None, None, None)
            assert 0 <= st <= 9
            if st == capi.RB_OK:
                ok += 1
                if descs[i].mix_start == capi.RB_MIX_START_CONSUMED:
                    assert out_len.value == 0
    assert ok > 100      # the generator does produce valid arrays too


def test_enum_values_agree_across_header_mirrors_and_oracle(built):
    """RB_FX_* / RB_FMT_* / RB_SIGNAL_* / flag values: include/rodio_b200.h is the definition; the ctypes mirror, the Rust binding's
    constants and the oracle's own enum (declared independently) must say the same numbers."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "rodio_b200.h")).read()
    values = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(RB_(?:FX|FMT|SIGNAL)_[A-Z0-9_]+)\s*=\s*(\d+)", hdr)}
    flags = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"\b(RB_[A-Z_]+)\s*=\s*1u\s*<<\s*(\d+)", hdr)}
    assert len(values) >= 29 and "RB_FX_PAUSE" in values and "RB_MIX_EXACT_ORDER" in flags
    for name, v in {**values, **flags}.items():
        if hasattr(capi, name):
            assert getattr(capi, name) == v, name
    for name in [n for n in values if n.startswith("RB_FX_")]:
        assert hasattr(capi, name), f"{name} missing from rodio_b200/_capi.py"
    rust = open(os.path.join(root, "bindings", "rust", "src", "lib.rs")).read()
    for m in re.finditer(r"pub const (RB_FX_[A-Z_]+): u32 = (\d+);", rust):
        assert values[m.group(1)] == int(m.group(2)), m.group(1)
    cap = open(os.path.join(root, "oracle", "rodio_oracle_capi.cpp")).read()
    m = re.search(r"enum \{ FX_AMPLIFY = 1,([^}]*)\}", cap)
    names = ["FX_AMPLIFY"] + [t.strip() for t in m.group(1).replace("\n", " ").split(",") if t.strip()]
    for i, n in enumerate(names, start=1):
        assert values["RB_" + n] == i, n

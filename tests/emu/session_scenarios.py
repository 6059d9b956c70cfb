"""Scenarios for tests/test_session_hostemu.py, run in a subprocess: the Python mirror (rodio_b200.Session) on top of
tests/emu/librodio_b200_hostemu.so -- the library's real host code (rb_api.cu) over the mock CUDA runtime, kernels on the SIMT
emulator.  The library path is swapped HERE, in the test process only; the product loader knows nothing of it.
    python tests/emu/session_scenarios.py <scenario> ...        exit code 0 = every scenario held bit for bit"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import build_emu                                      # noqa: E402
import rodio_b200._capi as capi                       # noqa: E402
capi.LIB_PATH = build_emu.HOST_LIB                    # RB_EMU_VARIANT picks a geometry variant of the kernel
import oracle                                         # noqa: E402
import rodio_b200 as rb                               # noqa: E402
from helpers import assert_bit_exact, assert_close_peak, noise, to_oracle     # noqa: E402
from test_lanes_emulator import expected_mix_classes                          # noqa: E402
import math                                           # noqa: E402

HELD = capi.RB_SESSION_HELD


def chain(pcm, ch_in, rate, mix_ch, mix_rate, lp, gain, speed=None, pre=None):
    s = rb.TestSource(pcm, ch_in, rate)
    if pre is not None:
        s = s.amplify(pre)           # source.amplify(v), then handed to the mixer: the gain is in front of the conversion
    if speed:
        s = s.speed(speed)
    s = rb.UniformSourceIterator(s, mix_ch, mix_rate)
    if lp:
        s = s.low_pass(lp)
    if gain is not None:
        s = s.amplify(gain)
    return s


def expected(pcms, ch_in, rates, mix_ch, mix_rate, joined, lp, gain, speeds=None, pres=None):
    speeds = speeds or [None] * len(pcms)
    pres = pres or [None] * len(pcms)
    srcs = [chain(p, ci, r, mix_ch, mix_rate, lp, gain, sp, pr) for p, ci, r, sp, pr in zip(pcms, ch_in, rates, speeds, pres)]
    per = [oracle.chain_uniform(to_oracle(s), mix_ch, mix_rate) for s in srcs]
    eff = [s.base_rate if sp is None else rb.capi.lib().rb_speed_sample_rate(r, sp) for s, r, sp in zip(srcs, rates, speeds)]
    froms = [r // math.gcd(r, mix_rate) for r in eff]
    tos = [mix_rate // math.gcd(r, mix_rate) for r in eff]
    n = max(j * mix_ch + y.size for j, y in zip(joined, per))
    return expected_mix_classes(per, [j * mix_ch for j in joined], n, froms, list(zip(tos, ch_in)))


def drive(sess, pcms, ch_in, block_frames, render_frames, rng=None, packed=True, hooks=None):
    """Push `block_frames[i]` frames of every source per round, render until nothing comes; hooks: {round: callable(sess)}."""
    pos, got, ended, rnd = [0] * len(pcms), [], False, 0
    while not ended:
        if hooks and rnd in hooks:
            r = hooks[rnd](sess)
            if r is not None:
                sess = r
        blocks, eos = [], []
        for i, (p, ci) in enumerate(zip(pcms, ch_in)):
            k = block_frames[i] if rng is None else int(rng.integers(0, 2 * block_frames[i] + 1))
            k = min(k, p.size // ci - pos[i])
            blocks.append(p[ci * pos[i]: ci * (pos[i] + k)])
            pos[i] += k
            eos.append(pos[i] == p.size // ci)
        if packed:
            sess.push_packed(blocks, eos)
        else:
            for i, (b, e) in enumerate(zip(blocks, eos)):
                if b.size or e:
                    sess.push(i, b, end_of_stream=e)
        while True:
            n = render_frames if rng is None else int(rng.integers(1, render_frames + 1))
            block, ended = sess.render(n)
            got.append(block)
            if block.size == 0 or ended:
                break
        rnd += 1
        assert rnd < 100000
    return np.concatenate(got), sess


def s_mono_random_split():
    rng = np.random.default_rng(1)
    pcms = [noise(int(n), 60 + i) for i, n in enumerate(rng.integers(300, 1500, 37))]
    ch = [1] * 37
    with rb.Session([chain(np.zeros(0, np.float32), 1, 44100, 1, 48000, 200, 1.2) for _ in pcms], 48000, fifo_frames=2048,
                    max_block_frames=500) as s:
        got, _ = drive(s, pcms, ch, [120] * 37, 500, rng=rng, packed=False)
    assert_bit_exact(got, expected(pcms, ch, [44100] * 37, 1, 48000, [0] * 37, 200, 1.2), "mono random split, single pushes")


def s_mixed_everything_with_state_blob():
    """Mono and stereo sources at three rates in a stereo mixer, packed 10 ms pushes, the state handed to a fresh session
    half-way (rb_session_get_state / set_state)."""
    ch_in = [1, 2, 1, 1, 2, 1, 2, 1] * 5
    rates = [44100, 44100, 48000, 22050, 48000, 44100, 44100, 48000] * 5
    pcms = [noise(ci * (int(0.025 * r) + 3 * i), 1500 + i, 0.8) for i, (ci, r) in enumerate(zip(ch_in, rates))]
    mk = lambda: [chain(np.zeros(0, np.float32), ci, r, 2, 48000, 800, 0.7) for ci, r in zip(ch_in, rates)]
    sa = rb.Session(mk(), 48000, fifo_frames=1024, max_block_frames=480, mixer_channels=2)
    sb = rb.Session(mk(), 48000, fifo_frames=1024, max_block_frames=480, mixer_channels=2)

    def hand_over(s):
        sb.set_state(s.get_state())
        return sb
    got, _ = drive(sa, pcms, ch_in, [r // 100 for r in rates], 480, hooks={2: hand_over})
    sa.close(), sb.close()
    assert_bit_exact(got, expected(pcms, ch_in, rates, 2, 48000, [0] * len(pcms), 800, 0.7), "mixed sources, blob hand-over")


def s_held_queue_gain_speed():
    """A queue of three sounds (one of them with speed(0.9)), a voice added later, a gain change, late mix_start."""
    rates = [44100, 48000, 22050, 44100, 32000]
    speeds = [None, None, 0.9, None, None]
    pcms = [noise(int(0.04 * r) + 7 * i, 3300 + i) for i, r in enumerate(rates)]
    ch = [1] * 5
    srcs = [chain(np.zeros(0, np.float32), 1, r, 1, 48000, 900, 0.8, sp) for r, sp in zip(rates, speeds)]
    starts = [0, HELD, HELD, HELD, 333]                      # 0 -> 1 -> 2 queue; 3 is added by hand; 4 is scheduled
    marks = {}
    with rb.Session(srcs, 48000, fifo_frames=8192, max_block_frames=256, mix_starts=starts) as s:
        s.follow(1, 0)
        s.follow(2, 1)

        def add_voice(sess):
            marks["T3"] = sum(b.size for b in marks["got"])   # frames rendered so far = where source 3 joins
            sess.start(3)
        pos, got, ended, rnd = [0] * 5, [], False, 0
        marks["got"] = got
        while not ended:
            if rnd == 4:
                add_voice(s)
            blocks, eos = [], []
            for i, p in enumerate(pcms):
                k = min(rates[i] // 100, p.size - pos[i])
                blocks.append(p[pos[i]:pos[i] + k])
                pos[i] += k
                eos.append(pos[i] == p.size)
            s.push_packed(blocks, eos)
            while True:
                block, ended = s.render(256)
                got.append(block)
                if block.size == 0 or ended:
                    break
            rnd += 1
    got = np.concatenate(got)
    lens = [oracle.chain_uniform(to_oracle(chain(p, 1, r, 1, 48000, 900, 0.8, sp)), 1, 48000).size for p, r, sp in zip(pcms, rates, speeds)]
    joined = [0, lens[0], lens[0] + lens[1], marks["T3"], 333]
    assert_bit_exact(got, expected(pcms, ch, rates, 1, 48000, joined, 900, 0.8, speeds), "queue + added voice + scheduled source")


def s_follow_after_predecessor_played_out():
    """Player::append on a player whose queue has run dry while another voice keeps the timeline moving: the sound starts at the
    current position (the frame rendered next), it neither stalls the session nor loses its head."""
    rates = [44100, 44100, 48000]
    pcms = [noise(100, 4100), noise(6000, 4101), noise(900, 4102)]
    ch = [1, 1, 1]
    srcs = [chain(np.zeros(0, np.float32), 1, r, 1, 48000, 700, 0.9) for r in rates]
    with rb.Session(srcs, 48000, fifo_frames=8192, max_block_frames=256, mix_starts=[0, 0, HELD]) as s:
        got = []
        s.push(0, pcms[0], end_of_stream=True)
        s.push(1, pcms[1][:1500])
        while True:
            block, ended = s.render(256)
            if block.size == 0:
                break
            got.append(block)
        T = sum(b.size for b in got)
        assert T > 200, T                      # source 0 (about 109 frames) has played out long ago
        s.follow(2, 0)
        s.push(2, pcms[2], end_of_stream=True)
        s.push(1, pcms[1][1500:], end_of_stream=True)
        frames, ended = s.available()
        assert frames > 0 and not ended, (frames, ended)
        while True:
            block, ended = s.render(256)
            got.append(block)
            if block.size == 0 or ended:
                break
    got = np.concatenate(got)
    assert_bit_exact(got, expected(pcms, ch, rates, 1, 48000, [0, 0, T], 700, 0.9), "follow after the predecessor has played out")


def s_held_and_queued_across_state_blob():
    """get_state / set_state carries `held` and `follows`: a started source keeps playing, a queued one keeps its place, a
    source that is still held stays held."""
    rates = [44100, 48000, 44100, 22050]
    pcms = [noise(1800 + 100 * i, 4200 + i) for i in range(4)]
    ch = [1] * 4
    mk = lambda: [chain(np.zeros(0, np.float32), 1, r, 1, 48000, 500, 1.1) for r in rates]
    starts = [0, HELD, HELD, HELD]            # 1 follows 0; 2 is started by hand before the hand-over; 3 after it
    sa = rb.Session(mk(), 48000, fifo_frames=8192, max_block_frames=200, mix_starts=starts)
    sb = rb.Session(mk(), 48000, fifo_frames=8192, max_block_frames=200, mix_starts=starts)
    sa.follow(1, 0)
    marks, got, pos, sess, rnd, ended = {}, [], [0] * 4, sa, 0, False
    while not ended:
        if rnd == 2:
            marks[2] = sum(b.size for b in got)
            sess.start(2)
        if rnd == 3:
            sb.set_state(sess.get_state())
            sess = sb
        if rnd == 5:
            marks[3] = sum(b.size for b in got)
            sess.start(3)
        blocks, eos = [], []
        for i, p in enumerate(pcms):
            k = min(rates[i] // 100, p.size - pos[i])
            blocks.append(p[pos[i]:pos[i] + k])
            pos[i] += k
            eos.append(pos[i] == p.size)
        sess.push_packed(blocks, eos)
        while True:
            block, ended = sess.render(200)
            got.append(block)
            if block.size == 0 or ended:
                break
        rnd += 1
        assert rnd < 10000
    sa.close(), sb.close()
    got = np.concatenate(got)
    len0 = oracle.chain_uniform(to_oracle(chain(pcms[0], 1, rates[0], 1, 48000, 500, 1.1)), 1, 48000).size
    assert_bit_exact(got, expected(pcms, ch, rates, 1, 48000, [0, len0, marks[2], marks[3]], 500, 1.1), "held / queued sources across a hand-over")


def s_skip_one():
    """Player::skip_one / stop (rb_session_skip): the current sound ends with the last frame the mixer's converter has pulled -- the
    converter still emits what it owes (the last frame raw) --, the queued one starts right behind it, another voice plays on."""
    rates = [44100, 32000, 48000]
    pcms = [noise(6000, 4300), noise(1500, 4301), noise(5000, 4302)]
    ch = [1, 1, 1]
    srcs = [chain(np.zeros(0, np.float32), 1, r, 1, 48000, 800, 0.9) for r in rates]
    with rb.Session(srcs, 48000, fifo_frames=8192, max_block_frames=256, mix_starts=[0, HELD, 0]) as s:
        s.follow(1, 0)
        got = []
        s.push(0, pcms[0][:2500])
        s.push(1, pcms[1], end_of_stream=True)
        s.push(2, pcms[2][:2600])
        for _ in range(6):
            block, _ = s.render(256)
            got.append(block)
        n0 = sum(b.size for b in got)                   # outputs of source 0 rendered so far (it started at 0)
        assert n0 > 0
        s.skip(0)
        fpos = min(((n0 - 1) * 147) // 160 + 2, 2500)   # frames the converter had pulled: its whole input now
        s.push(2, pcms[2][2600:], end_of_stream=True)
        try:
            s.push(0, pcms[0][2500:])
        except rb.RodioB200Error:
            pass
        else:
            raise AssertionError("a push into a skipped source was accepted")
        while True:
            block, ended = s.render(256)
            got.append(block)
            if block.size == 0 or ended:
                break
    got = np.concatenate(got)
    cut = [pcms[0][:fpos], pcms[1], pcms[2]]
    len0 = oracle.chain_uniform(to_oracle(chain(cut[0], 1, rates[0], 1, 48000, 800, 0.9)), 1, 48000).size
    assert_bit_exact(got, expected(cut, ch, rates, 1, 48000, [0, len0, 0], 800, 0.9), "skip_one: the sound ends where the converter stood")
    # a held source that is skipped never plays
    with rb.Session(srcs[:2], 48000, fifo_frames=4096, max_block_frames=256, mix_starts=[0, HELD]) as s:
        s.push(0, pcms[1], end_of_stream=True)
        s.push(1, pcms[2][:1000], end_of_stream=True)
        s.skip(1)
        got = []
        while True:
            block, ended = s.render(256)
            got.append(block)
            if block.size == 0 or ended:
                break
    assert_bit_exact(np.concatenate(got), expected([pcms[1]], [1], [rates[0]], 1, 48000, [0], 800, 0.9), "a skipped held source never plays")


def s_filtered_and_plain_sources():
    """A low-passed source beside an untouched one (examples/stream_mixer.c): classes of their own, two launches."""
    pcms = [noise(2 * 900, 51), noise(1000, 52), noise(2 * 700, 53), noise(800, 54)]
    ch_in, rates = [2, 1, 2, 1], [44100, 48000, 48000, 22050]
    srcs = [chain(np.zeros(0, np.float32), 2, 44100, 2, 48000, 200, 0.8), chain(np.zeros(0, np.float32), 1, 48000, 2, 48000, None, None),
            chain(np.zeros(0, np.float32), 2, 48000, 2, 48000, None, 0.5), chain(np.zeros(0, np.float32), 1, 22050, 2, 48000, 300, None)]
    with rb.Session(srcs, 48000, fifo_frames=2048, max_block_frames=480, mixer_channels=2) as s:
        got, _ = drive(s, pcms, ch_in, [r // 100 for r in rates], 480)
    real = [chain(pcms[0], 2, 44100, 2, 48000, 200, 0.8), chain(pcms[1], 1, 48000, 2, 48000, None, None),
            chain(pcms[2], 2, 48000, 2, 48000, None, 0.5), chain(pcms[3], 1, 22050, 2, 48000, 300, None)]
    per = [oracle.chain_uniform(to_oracle(x), 2, 48000) for x in real]
    # every source is a class of its own here (rate pair x channels x filtered): the mix is the plain sum in source order
    acc = np.zeros(max(y.size for y in per), np.float32)
    for y in per:
        row = np.zeros(acc.size, np.float32)
        row[:y.size] = y
        acc = acc + (row + np.float32(0.0))
    assert_bit_exact(got, acc, "filtered beside plain sources")
    ref = oracle.mixer([to_oracle(x) for x in real], 2, 48000)
    assert_close_peak(got, ref, 1e-5, "... and the reference's mixer")


def s_gain_in_front_of_the_conversion():
    """`source.amplify(v)` handed to the mixer -- the usual rodio idiom -- in a session (random split, one source sped up, one
    gain outside the range of the fast tiles) and in a batch on the lane kernel."""
    rng = np.random.default_rng(9)
    n = 36
    rates = [44100, 48000, 22050, 44100] * 9
    ch_in = [2, 1, 1, 2] * 9
    pres = [float(np.float32(v)) for v in rng.uniform(0.1, 1.2, n)]
    pres[3], pres[10] = 0.004, -0.7
    speeds = [None] * n
    speeds[6] = 1.25
    pcms = [noise(ci * (500 + 17 * i), 8100 + i) for i, ci in enumerate(ch_in)]
    starts = [0 if i % 3 else 20 * i for i in range(n)]
    srcs = [chain(np.zeros(0, np.float32), ci, r, 2, 48000, 400, 0.9, sp, pr) for ci, r, sp, pr in zip(ch_in, rates, speeds, pres)]
    with rb.Session(srcs, 48000, fifo_frames=2048, max_block_frames=256, mix_starts=starts, mixer_channels=2) as s:
        got, _ = drive(s, pcms, ch_in, [r // 200 for r in rates], 256, rng=rng)
    want = expected(pcms, ch_in, rates, 2, 48000, starts, 400, 0.9, speeds, pres)
    assert_bit_exact(got, want[:got.size], "session with gains in front of the conversion")
    assert got.size == want.size or not np.any(want[got.size:])
    # sources with and without such a gain in one session: classes of their own, summed in class order
    mixed = [chain(np.zeros(0, np.float32), 1, 44100, 1, 48000, None, None, None, 0.5), chain(np.zeros(0, np.float32), 1, 44100, 1, 48000, None, None)]
    with rb.Session(mixed, 48000, fifo_frames=1024, max_block_frames=128, mixer_channels=1) as s:
        got, _ = drive(s, [pcms[1][:500], pcms[2][:500]], [1, 1], [100, 100], 128)
    a = oracle.chain_uniform(to_oracle(chain(pcms[1][:500], 1, 44100, 1, 48000, None, None, None, 0.5)), 1, 48000)
    b = oracle.chain_uniform(to_oracle(chain(pcms[2][:500], 1, 44100, 1, 48000, None, None)), 1, 48000)
    assert_bit_exact(got, (a + np.float32(0.0)) + (b + np.float32(0.0)), "with and without a gain in front")
    # the same shape as a batch: rb_batch_create -> fused parser -> lane plan
    real = [chain(p, ci, r, 2, 48000, 400, 0.9, None, pr) for p, ci, r, pr in zip(pcms, ch_in, [44100, 22050] * 18, pres)]
    with rb.Batch(real, 2, 48000, flags=capi.RB_FUSED_LANES) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
    assert_bit_exact(got, expected(pcms, ch_in, [44100, 22050] * 18, 2, 48000, [0] * n, 400, 0.9, None, pres), "batch with gains in front")


def front_chain(pcm, ch_in, rate, mix_ch, mix_rate, lp, pre=None, mid=None, gain=None, speed=None):
    """A Player's chain with a user filter: source [.speed] [.amplify(pre)] .low_pass(lp) [.amplify(mid)] -> mixer conversion [-> amplify]."""
    s = rb.TestSource(pcm, ch_in, rate)
    if speed:
        s = s.speed(speed)
    if pre is not None:
        s = s.amplify(pre)
    s = s.low_pass(lp)
    if mid is not None:
        s = s.amplify(mid)
    s = rb.UniformSourceIterator(s, mix_ch, mix_rate)
    if gain is not None:
        s = s.amplify(gain)
    return s


def s_filter_in_front_of_the_conversion():
    """`source.low_pass(f)` handed to the mixer / appended to a Player (player.rs:120-128: the volume sits behind it): the filter
    runs at the source's rate.  A session with random splits (up- and down-sampling sources, one sped up, mono in stereo), its
    state handed to a second session half way, and the same shape as a batch on the lane kernel."""
    rng = np.random.default_rng(21)
    n = 30
    rates = [44100, 48000, 22050, 96000, 44100, 32000] * 5
    ch_in = [2, 1, 1, 2, 1, 2] * 5
    mids = [float(np.float32(v)) for v in rng.uniform(0.2, 1.1, n)]
    speeds = [None] * n
    speeds[4] = 1.1
    pcms = [noise(ci * (600 + 23 * i), 9100 + i) for i, ci in enumerate(ch_in)]
    starts = [0 if i % 4 else 15 * i for i in range(n)]
    mk = lambda p, i: front_chain(p, ch_in[i], rates[i], 2, 48000, 350, 0.9, mids[i], 0.8, speeds[i])
    srcs = [mk(np.zeros(0, np.float32), i) for i in range(n)]
    real = [mk(pcms[i], i) for i in range(n)]
    per = [oracle.chain_uniform(to_oracle(x), 2, 48000) for x in real]
    eff = [rates[i] if speeds[i] is None else rb.capi.lib().rb_speed_sample_rate(rates[i], speeds[i]) for i in range(n)]
    froms = [r // math.gcd(r, 48000) for r in eff]
    tos = [48000 // math.gcd(r, 48000) for r in eff]
    want = expected_mix_classes(per, [2 * j for j in starts], max(2 * j + y.size for j, y in zip(starts, per)), froms, list(zip(tos, ch_in)))

    def hand_over(s):
        blob = s.get_state()
        t = rb.Session(srcs, 48000, fifo_frames=4096, max_block_frames=256, mix_starts=starts, mixer_channels=2)
        t.set_state(blob)
        s.close()
        return t
    s = rb.Session(srcs, 48000, fifo_frames=4096, max_block_frames=256, mix_starts=starts, mixer_channels=2)
    got, s = drive(s, pcms, ch_in, [r // 200 for r in rates], 256, rng=rng, hooks={3: hand_over})
    s.close()
    assert_bit_exact(got, want[:got.size], "session with the filter in front of the conversion")
    assert got.size == want.size or not np.any(want[got.size:])
    # the same as a batch (up-sampling and same-rate sources: the planner drops the identity conversion of the 48 kHz ones)
    rates_b = [44100, 48000, 22050] * 10
    real = [front_chain(pcms[i], ch_in[i], rates_b[i], 2, 48000, 350, 0.9, mids[i], 0.8) for i in range(n)]
    with rb.Batch(real, 2, 48000, flags=capi.RB_FUSED_LANES) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
    per = [oracle.chain_uniform(to_oracle(x), 2, 48000) for x in real]
    froms = [r // math.gcd(r, 48000) for r in rates_b]
    tos = [48000 // math.gcd(r, 48000) for r in rates_b]
    assert_bit_exact(got, expected_mix_classes(per, [0] * n, max(y.size for y in per), froms, list(zip(tos, ch_in))), "batch with the filter in front")
    # a filter in front AND one behind: not a session shape
    both = rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 1, 44100).low_pass(300), 1, 48000).low_pass(500)
    try:
        rb.Session([both], 48000, mixer_channels=1)
    except capi.RodioB200Error as e:
        assert e.status == capi.RB_ERR_UNSUPPORTED
    else:
        raise AssertionError("two filters were accepted")


def s_player_volume_changes():
    """Player::set_volume while playing (rb_session_set_volume): the Player's Amplify sits in front of the mixer's conversion
    (behind the user's filter), so a frame keeps the factor it had when the converter pulled it -- frames pulled in a block
    carry that block's factor, the two look-ahead frames of the converter included.  Expectation: the oracle's literal chain
    over an input that was multiplied frame by frame with exactly those factors."""
    rates = [44100, 22050, 48000, 96000, 88200, 88200]
    front = [False, True, False, True, True, False]
    L = [6000, 3000, 6600, 13000, 40000, 12100]
    pcms = [noise(n, 8800 + i) for i, n in enumerate(L)]
    mk_plain = lambda p, r, g: rb.UniformSourceIterator(rb.TestSource(p, 1, r).amplify(g), 1, 48000).low_pass(500).amplify(0.9)
    mk_front = lambda p, r, g: rb.UniformSourceIterator(rb.TestSource(p, 1, r).amplify(0.8).low_pass(400).amplify(g), 1, 48000)
    srcs = [(mk_front if f else mk_plain)(np.zeros(0, np.float32), r, 1.0) for r, f in zip(rates, front)]
    plan = {2: (0, 0.5), 3: (1, 0.25), 5: (0, 1.5), 6: (3, 0.1), 7: (2, 0.7), 8: (0, 0.0), 9: (1, 1.0), 10: (3, 0.9), 11: (0, 0.6),
            12: (1, 0.0), 4: (4, 0.3), 13: (4, 1.2), 1: (5, 0.45), 14: (5, 0.9)}   # round -> (stream, volume); 0.0 = muted
    vol = [1.0] * len(rates)
    busy = {r: (4, 0.3 + 0.04 * r) for r in range(15, 40)}      # and one source whose volume moves in every round for a while
    plan = {**busy, **plan}
    gains = [np.ones(n, np.float32) for n in L]       # factor of every input frame
    pulled, pushed, done = [0] * len(rates), [0] * len(rates), [0] * len(rates)
    total = [int(rb.plan((mk_front if f else mk_plain)(p, r, 1.0), 1, 48000)[0]) for p, r, f in zip(pcms, rates, front)]
    got, ended, rnd = [], False, 0
    with rb.Session(srcs, 48000, fifo_frames=4096, max_block_frames=240, mixer_channels=1) as s:
        while not ended:
            if rnd == 4:      # hand the session over: the blob carries the factors and those of the look-ahead frames
                t = rb.Session(srcs, 48000, fifo_frames=4096, max_block_frames=240, mixer_channels=1)
                t.set_state(s.get_state())
                s.close()
                s = t
            if rnd in plan:
                i, v = plan[rnd]
                s.set_volume(i, v)
                vol[i] = v
            blocks, eos = [], []
            for i, (p, r) in enumerate(zip(pcms, rates)):
                k = min(r // 100, L[i] - pushed[i])
                blocks.append(p[pushed[i]:pushed[i] + k])
                pushed[i] += k
                eos.append(pushed[i] == L[i])
            s.push_packed(blocks, eos)
            while True:
                block, ended = s.render(240)
                got.append(block)
                if block.size == 0:
                    break
                for i, r in enumerate(rates):          # what the converter of every source pulled in this block
                    o = min(done[i] + block.size, total[i])
                    if o > done[i]:
                        g = math.gcd(r, 48000)
                        p_new = min(((o - 1) * (r // g)) // (48000 // g) + 2, pushed[i])
                        gains[i][pulled[i]:p_new] = np.float32(vol[i])
                        pulled[i], done[i] = p_new, o
                if ended:
                    break
            rnd += 1
            assert rnd < 1000
        s.close()         # the session that took over
    got = np.concatenate(got)
    per = []
    for i, (p, r, f) in enumerate(zip(pcms, rates, front)):
        if f:     # amplify(0.8) -> low_pass at the source's rate, then the Player's factor on every filter output
            y = oracle.chain_uniform(to_oracle(rb.TestSource(p, 1, r).amplify(0.8).low_pass(400)), 1, r)
            per.append(oracle.chain_uniform(to_oracle(rb.UniformSourceIterator(rb.TestSource(y * gains[i], 1, r), 1, 48000)), 1, 48000))
        else:
            x = rb.UniformSourceIterator(rb.TestSource(p * gains[i], 1, r), 1, 48000).low_pass(500).amplify(0.9)
            per.append(oracle.chain_uniform(to_oracle(x), 1, 48000))
    acc = np.zeros(max(y.size for y in per), np.float32)
    for y in per:                                      # every source is a class of its own: the plain sum in source order
        row = np.zeros(acc.size, np.float32)
        row[:y.size] = y
        acc = acc + (row + np.float32(0.0))
    assert all(np.unique(g).size > 1 for g in gains), "every source saw a volume change"
    assert_bit_exact(got, acc, "volume changes in front of the conversion")
    # a source declared without an AMPLIFY in front cannot get one later
    with rb.Session([rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 1, 44100), 1, 48000)], 48000, mixer_channels=1) as s:
        try:
            s.set_volume(0, 0.5)
        except capi.RodioB200Error as e:
            assert e.status == capi.RB_ERR_STATE
        else:
            raise AssertionError("set_volume without an AMPLIFY in front was accepted")


def s_batch_with_identity_conversions():
    """rb_batch_create -> fused parser -> lane plan on the CPU: 44.1 kHz sources beside 48 kHz ones (whose conversion is the
    identity and is dropped by the planner) and mono beside stereo -- the batch still goes to the lane kernel, class by class."""
    rates = [44100, 48000, 22050, 48000, 44100, 48000] * 2
    ch_in = [1, 2, 1, 1, 2, 2] * 2
    pcms = [noise(ci * (400 + 9 * i), 4500 + i) for i, ci in enumerate(ch_in)]
    srcs = [chain(p, ci, r, 2, 48000, 500, 0.9) for p, ci, r in zip(pcms, ch_in, rates)]
    with rb.Batch(srcs, 2, 48000, flags=capi.RB_FUSED_LANES) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
        assert np.array_equal(got.view(np.uint32), b.render_mix().view(np.uint32))
    assert_bit_exact(got, expected(pcms, ch_in, rates, 2, 48000, [0] * len(pcms), 500, 0.9), "batch with identity conversions")
    # a chain with a gain in front of the filter files it differently for the two row kinds: not the lane kernel's shape
    odd = [rb.UniformSourceIterator(rb.TestSource(pcms[0], 1, 44100), 2, 48000).amplify(0.5).low_pass(300),
           rb.UniformSourceIterator(rb.TestSource(pcms[3], 1, 48000), 2, 48000).amplify(0.5).low_pass(300)]
    with rb.Batch(odd, 2, 48000, flags=capi.RB_FUSED_LANES) as b:
        assert b.kernel_family != 2


def s_batch_unsorted_starts():
    """Mixer::add calls in arbitrary timeline order: the batch orders its sources by mix_start (stable) -- that order is the
    lane order; the rendered mix is also served in blocks (rb_batch_read_mix)."""
    from helpers import lanes_expected_mix
    rng = np.random.default_rng(4)
    n = 40
    pcms = [noise(300 + 11 * i, 7000 + i) for i in range(n)]
    starts = [int(v) for v in rng.integers(0, 400, n)]
    srcs = [chain(p, 1, 44100, 1, 48000, 700, 0.9) for p in pcms]
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_FUSED_LANES, mix_starts=starts) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
        blocks = np.concatenate([b.read_mix(o, 100) for o in range(0, got.size, 100)])
    assert np.array_equal(blocks.view(np.uint32), got.view(np.uint32))
    order = sorted(range(n), key=lambda i: starts[i])
    per = [oracle.chain_uniform(to_oracle(srcs[i]), 1, 48000) for i in order]
    assert_bit_exact(got, lanes_expected_mix(per, [starts[i] for i in order], got.size), "unsorted mix_starts")


def s_gain_changes():
    """rb_session_set_amplify between 5 ms blocks, rb_session_available in step with what render delivers."""
    pcms = [noise(1500, 800 + i) for i in range(3)]
    srcs = [chain(np.zeros(0, np.float32), 1, 44100, 1, 48000, 500, 1.0) for _ in pcms]
    per = [oracle.chain_uniform(to_oracle(chain(p, 1, 44100, 1, 48000, 500, 1.0)), 1, 48000) for p in pcms]
    rng = np.random.default_rng(3)
    got, gains, block = [], [], 240
    with rb.Session(srcs, 48000, fifo_frames=4096, max_block_frames=block) as s:
        for i, p in enumerate(pcms):
            s.push(i, p, end_of_stream=True)
        ended = False
        while not ended:
            g = [np.float32(v) for v in rng.uniform(0.0, 1.5, 3)]
            for i in range(3):
                s.set_amplify(i, float(g[i]))
            avail, _ = s.available()
            out, ended = s.render(block)
            assert out.size == min(avail, block)
            if out.size:
                got.append(out), gains.append(g)
    got = np.concatenate(got)
    scaled = []
    for r in range(3):
        y = per[r].copy()
        for k, g in enumerate(gains):
            y[k * block:(k + 1) * block] = y[k * block:(k + 1) * block] * g[r]
        scaled.append(y)
    assert_bit_exact(got, expected_mix_classes(scaled, [0] * 3, got.size, [147] * 3, [(160, 1)] * 3), "per-block gains")


def s_duo_batches():
    """The lane-pair kernel (rb_duo_core.h, RB_FUSED_DUO) through the library's own planner: bit for bit against the oracle's
    streams added in the kernel's documented order (pairs, then the tree over 64), within the tolerance of the sequential sum."""
    from helpers import duo_expected_mix, lanes_expected_mix

    def run(srcs, starts, family, what):
        with rb.Batch(srcs, 1, 48000, flags=capi.RB_FUSED_DUO, mix_starts=starts) as b:
            assert b.kernel_family == family, (what, b.kernel_family)
            assert b.mix_group == (64 if family == 3 else 32)
            b.upload_all()
            got = b.render_mix()
            again = b.render_mix()
        assert np.array_equal(got.view(np.uint32), again.view(np.uint32)), what
        per = [oracle.chain_uniform(to_oracle(x), 1, 48000) for x in srcs]
        ref = oracle.mixer([to_oracle(x, st) for x, st in zip(srcs, starts)], 1, 48000)
        assert_close_peak(got, ref, 1e-5, what + ": vs the sequential mixer")
        want = (duo_expected_mix if family == 3 else lanes_expected_mix)(per, starts, ref.size)
        assert_bit_exact(got, want, what)
    rng = np.random.default_rng(77)
    # cfg3 shape, more than one group, an odd number of streams (the last lane has one row, the last group is partial)
    pcms = [noise(1500 + 7 * i, 8000 + i) for i in range(131)]
    run([chain(p, 1, 44100, 1, 48000, 200, 1.2) for p in pcms], [0] * 131, 3, "cfg3 shape")
    # ragged lengths (empty, one frame, shorter than a tile), starts that differ by multiples of 4 * to = 640 frames
    lens = [0, 1, 2, 7, 37, 900, 2500, 16, 17, 640, 641, 3000] + [int(v) for v in rng.integers(3, 2500, 40)]
    starts = sorted(640 * int(v) for v in rng.integers(0, 4, len(lens)))
    run([chain(noise(n, 8200 + i), 1, 44100, 1, 48000, 1000, 0.7) for i, n in enumerate(lens)], starts, 3, "ragged, starts in phase")
    # high_pass (ffk = -2), no gain; then no filter at all
    run([rb.UniformSourceIterator(rb.TestSource(noise(1200 + i, 8300 + i), 1, 22050), 1, 48000).high_pass(300) for i in range(70)], [0] * 70, 3, "high_pass, 22.05 kHz")
    run([chain(noise(1000 + 3 * i, 8400 + i), 1, 44100, 1, 48000, None, 0.8) for i in range(66)], [0] * 66, 3, "no filter")
    run([chain(noise(1000 + 3 * i, 8500 + i), 1, 32000, 1, 48000, None, None) for i in range(10)], [0] * 10, 3, "nothing but the conversion")
    # inputs outside the exact-reciprocal class keep their lane on the slow tiles
    bad = [noise(900, 8600 + i) for i in range(40)]
    bad[3][100] = 1e-42
    bad[17][5] = 1e25
    bad[18][6] = -0.0
    run([chain(p, 1, 44100, 1, 48000, 500, 1.1) for p in bad], [0] * 40, 3, "unsafe inputs")
    # neighbours out of phase: the class stays on k_fused_lanes
    starts = sorted(int(v) for v in rng.integers(0, 300, 50))
    run([chain(noise(800 + i, 8700 + i), 1, 44100, 1, 48000, 200, 1.2) for i in range(50)], starts, 2, "starts out of phase")
    # several rate pairs: classes of their own, each on the pair kernel
    rates = [44100, 22050, 32000] * 30
    srcs = [chain(noise(700 + 5 * i, 8800 + i), 1, r, 1, 48000, 600, 0.9) for i, r in enumerate(rates)]
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_FUSED_DUO) as b:
        assert b.kernel_family == 3
        b.upload_all()
        got = b.render_mix()
    per = [oracle.chain_uniform(to_oracle(x), 1, 48000) for x in srcs]
    acc = np.zeros(got.size, np.float32)
    for r in (44100, 22050, 32000):
        idx = [i for i, q in enumerate(rates) if q == r]
        acc = acc + (duo_expected_mix([per[i] for i in idx], [0] * len(idx), got.size) - np.float32(0.0))
    assert_bit_exact(got, acc, "three rate pairs on the pair kernel")


def s_time_parallel_plan():
    """RB_BIQUAD_TIME_PARALLEL: low_pass(1000) batches run as timeline segments with a warm-up (family 4), within the north-star
    tolerance of the reference and no further from an f64 run of the same filter than the reference itself; low_pass(200) is
    refused by the accuracy gate and served exactly."""
    os.environ["RB_TP_SEGMENTS"] = "5"
    n, frames = 70, 9000
    pcms = [noise(frames - 40 * (i % 9), 9300 + i) for i in range(n)]
    starts = [0] * n
    starts[5], starts[6], starts[40] = 640, 1280, 3200     # late joiners, in phase (multiples of 4 * to)
    starts = sorted(starts)
    for lp, q, want_family in ((1000, 0.5, 4), (3000, 0.707, 4), (200, 0.5, 3)):
        srcs = [rb.UniformSourceIterator(rb.TestSource(p, 1, 44100), 1, 48000).low_pass_with_q(lp, q).amplify(1.2) for p in pcms]
        with rb.Batch(srcs, 1, 48000, flags=capi.RB_BIQUAD_TIME_PARALLEL | capi.RB_FUSED_DUO, mix_starts=starts) as b:
            assert b.kernel_family == want_family, (lp, b.kernel_family)
            b.upload_all()
            got = b.render_mix()
            again = b.render_mix()
        assert np.array_equal(got.view(np.uint32), again.view(np.uint32))
        ref = oracle.mixer([to_oracle(x, st) for x, st in zip(srcs, starts)], 1, 48000)
        assert_close_peak(got, ref, 1e-5, f"time-parallel low_pass({lp}) vs the reference")
        if want_family == 4:
            per = [oracle.chain_uniform(to_oracle(x), 1, 48000) for x in srcs]
            from helpers import duo_expected_mix
            exact = duo_expected_mix(per, starts, ref.size)
            same = float(np.mean(got.view(np.uint32) == exact.view(np.uint32)))
            print(f"low_pass({lp}, q={q}): {100 * same:.1f} % of the mix bit-identical to the serial run in the same order")
            assert same > 0.5
    # without the duo flag and below the automatic threshold the flag alone selects the plan
    srcs = [rb.UniformSourceIterator(rb.TestSource(p, 1, 44100), 1, 48000).low_pass(2000) for p in pcms]
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_BIQUAD_TIME_PARALLEL) as b:
        assert b.kernel_family == 4
        b.upload_all()
        got = b.render_mix()
    assert_close_peak(got, oracle.mixer([to_oracle(x) for x in srcs], 1, 48000), 1e-5, "time-parallel, no gain")
    del os.environ["RB_TP_SEGMENTS"]


def s_errors():
    mk = lambda ci, r, mc: chain(np.zeros(0, np.float32), ci, r, mc, 48000, 200, None)
    for bad in (lambda: rb.Session([mk(2, 44100, 1)], 48000, mixer_channels=1),                       # stereo source, mono mixer
                lambda: rb.Session([rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 1, 44100), 1, 48000)
                                    .automatic_gain_control()], 48000),                                 # adapter outside the shape
                lambda: rb.Session([rb.TestSource(np.zeros(0, np.float32), 1, 44100).low_pass(100)], 48000)):   # no UNIFORM
        try:
            bad()
        except rb.RodioB200Error:
            continue
        raise AssertionError("a session outside the served shape was accepted")
    with rb.Session([mk(1, 44100, 1)], 48000, fifo_frames=64, max_block_frames=64) as s:
        s.push(0, noise(64, 1))
        try:
            s.push(0, noise(1, 2))
        except rb.RodioB200Error as e:
            assert "FIFO" in str(e) or "full" in str(e).lower(), str(e)
        else:
            raise AssertionError("a push beyond the FIFO was accepted")
        out, ended = s.render(64)
        assert out.size > 0 and not ended
        s.push(0, noise(10, 3), end_of_stream=True)
        try:
            s.push(0, noise(1, 4))
        except rb.RodioB200Error:
            pass
        else:
            raise AssertionError("a push after end_of_stream was accepted")
    # a session whose FIFOs could not exist is refused before any allocation (no size_t wrap-around)
    big = [chain(np.zeros(0, np.float32), 1, 44100, 1, 48000, None, None) for _ in range(2000)]
    try:
        rb.Session(big, 48000, fifo_frames=2 ** 32 - 1, max_block_frames=480, mixer_channels=1)
    except capi.RodioB200Error as e:
        assert e.status == capi.RB_ERR_OUT_OF_MEMORY
    else:
        raise AssertionError("a 32 TB session was accepted")


def s_random(seed=0, cases=6):
    """Randomised sessions through the real host code (not part of the suite: python tests/emu/session_scenarios.py random:SEED)."""
    rng = np.random.default_rng(seed)
    for case in range(cases):
        mix_rate = int(rng.choice([48000, 44100]))
        mix_ch = int(rng.choice([1, 2]))
        n = int(rng.integers(1, 45))
        rates = [int(rng.choice([44100, 22050, 48000, 32000, 96000, 8000])) for _ in range(n)]
        ch_in = [mix_ch if (mix_ch == 1 or rng.random() < 0.6) else 1 for _ in range(n)]
        lens = [int(rng.integers(0, 1200)) for _ in range(n)]
        pcms = [noise(ci * L, 991 * case + i + 7 * seed) for i, (ci, L) in enumerate(zip(ch_in, lens))]
        starts = [0 if rng.random() < 0.6 else int(rng.integers(0, 500)) for _ in range(n)]
        lp, gain = [(300, 0.8), (None, 1.1), (2000, None), (None, None)][int(rng.integers(4))]
        pres = [float(np.float32(rng.choice([0.002, -0.4, rng.uniform(0.05, 1.5)]))) for _ in range(n)] if rng.random() < 0.4 else None
        front = lp is not None and rng.random() < 0.5
        if front:     # the filter in front of the conversion, the Player's volume behind it
            mids = [float(np.float32(rng.uniform(0.2, 1.2))) for _ in range(n)]
            mk = lambda p, i: front_chain(p, ch_in[i], rates[i], mix_ch, mix_rate, lp, None if pres is None else pres[i], mids[i], gain)
            srcs = [mk(np.zeros(0, np.float32), i) for i in range(n)]
            with rb.Session(srcs, mix_rate, fifo_frames=4096, max_block_frames=int(rng.choice([64, 333, 1024])), mix_starts=starts,
                            mixer_channels=mix_ch) as s:
                got, _ = drive(s, pcms, ch_in, [max(1, r // 150) for r in rates], 700, rng=rng, packed=bool(rng.integers(2)))
            per = [oracle.chain_uniform(to_oracle(mk(pcms[i], i)), mix_ch, mix_rate) for i in range(n)]
            froms = [r // math.gcd(r, mix_rate) for r in rates]
            tos = [mix_rate // math.gcd(r, mix_rate) for r in rates]
            want = expected_mix_classes(per, [j * mix_ch for j in starts], max(j * mix_ch + y.size for j, y in zip(starts, per)), froms,
                                        list(zip(tos, ch_in)))
            if any(lens):
                assert_bit_exact(got, want[:got.size], f"random session {seed}/{case} (filter in front)")
                assert got.size == want.size or not np.any(want[got.size:]), "tail"
            continue
        srcs = [chain(np.zeros(0, np.float32), ci, r, mix_ch, mix_rate, lp, gain, None, None if pres is None else pres[i])
                for i, (ci, r) in enumerate(zip(ch_in, rates))]
        with rb.Session(srcs, mix_rate, fifo_frames=4096, max_block_frames=int(rng.choice([64, 333, 1024])), mix_starts=starts,
                        mixer_channels=mix_ch) as s:
            got, _ = drive(s, pcms, ch_in, [max(1, r // 150) for r in rates], 700, rng=rng, packed=bool(rng.integers(2)))
        want = expected(pcms, ch_in, rates, mix_ch, mix_rate, starts, lp, gain, None, pres)
        if not any(lens):
            assert got.size == 0
            continue
        assert_bit_exact(got, want[:got.size], f"random session {seed}/{case}")
        assert got.size == want.size or not np.any(want[got.size:]), "tail"


SCENARIOS = {"mono_random_split": s_mono_random_split, "mixed_with_state_blob": s_mixed_everything_with_state_blob,
             "held_queue_gain_speed": s_held_queue_gain_speed, "follow_after_played_out": s_follow_after_predecessor_played_out,
             "held_across_state_blob": s_held_and_queued_across_state_blob, "skip_one": s_skip_one, "duo_batches": s_duo_batches, "time_parallel_plan": s_time_parallel_plan, "gain_changes": s_gain_changes, "filtered_and_plain": s_filtered_and_plain_sources,
             "batch_with_identity_conversions": s_batch_with_identity_conversions, "batch_unsorted_starts": s_batch_unsorted_starts, "errors": s_errors,
             "gain_in_front": s_gain_in_front_of_the_conversion, "filter_in_front": s_filter_in_front_of_the_conversion, "player_volume": s_player_volume_changes}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SCENARIOS)):
        if name.startswith("random"):
            s_random(int(name.split(":")[1]) if ":" in name else 0)
        else:
            SCENARIOS[name]()
        print("ok", name, flush=True)

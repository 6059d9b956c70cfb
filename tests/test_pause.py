"""Pausable (src/source/pausable.rs:8-21,:85-97) as a batch adapter, RB_FX_PAUSE: Player::pause / play scripted.  While paused the
adapters in front are NOT pulled (a filter there keeps its state), whole frames of zeros are emitted, then the source carries on."""
import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, noise, to_oracle


def _chain(src):
    return oracle.chain(to_oracle(src))[0]


def test_oracle_pause_freezes_the_filter_in_front():
    x = noise(2 * 3000, 5)
    filtered = _chain(rb.TestSource(x, 2, 44100).low_pass(300))
    for at, frames in ((0, 3), (1001, 50), (6000, 9), (2, 1)):
        got = _chain(rb.TestSource(x, 2, 44100).low_pass(300).pause_at(at, frames))
        want = np.concatenate([filtered[:at], np.zeros(2 * frames, np.float32), filtered[at:]])
        assert_bit_exact(got, want, f"pause at {at} for {frames} frames")
    # a pause behind the last sample never sets in; a pause of no frames is nothing
    assert_bit_exact(_chain(rb.TestSource(x, 2, 44100).pause_at(6001, 5)), x, "pause behind the end")
    assert_bit_exact(_chain(rb.TestSource(x, 2, 44100).pause_at(10, 0)), x, "pause of nothing")


@pytest.mark.gpu
def test_pause_bit_exact_on_the_device(ctx):
    x = noise(2 * 3000, 6)
    for at, frames in ((0, 3), (1001, 50), (6000, 9), (6001, 4)):
        src = rb.TestSource(x, 2, 44100).low_pass(300).pause_at(at, frames).amplify(0.8)
        assert_bit_exact(src.collect(ctx), _chain(src), f"pause at {at} for {frames} frames")
    srcs = [rb.TestSource(x, 2, 44100).high_pass(200).pause_at(2000, 480), rb.SineWave(330.0).take(4000).pause_at(1000, 100).amplify(0.5)]
    want = oracle.mixer([to_oracle(s) for s in srcs], 2, 48000)
    with rb.Batch(srcs, 2, 48000, flags=rb.capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "paused sources in a mixer")


@pytest.mark.gpu
def test_player_pause_and_play_at_known_positions(ctx):
    """Player::pause / play through the Player mirror: Pausable sits between the speed and the volume (player.rs:122-128), so the
    user's filter is frozen while paused and the silence is NOT scaled into anything else; two pauses in one sound."""
    x = noise(2 * 5000, 7)
    tx, rx = rb.mixer(2, 48000, ctx=ctx)
    player = rb.Player.connect_new(tx)
    player.set_volume(0.5)
    player.append(rb.TestSource(x, 2, 44100).low_pass(300), pauses=[(2000, 100), (6000, 30)])
    player.append(rb.SineWave(440.0).take(3000), pauses=[(0, 10)])
    got = rx.collect()
    first = rb.TestSource(x, 2, 44100).low_pass(300).speed(1.0).pause_at(2000, 100).pause_at(6000 + 200, 30).amplify(0.5)
    second = rb.SineWave(440.0).take(3000).speed(1.0).pause_at(0, 10).amplify(0.5)
    want = np.concatenate([oracle.chain_uniform(to_oracle(first), 2, 48000), oracle.chain_uniform(to_oracle(second), 2, 48000)]) + np.float32(0.0)
    assert_bit_exact(got, want, "player with pauses")
    # the filter did not run through the silence: behind the pause the samples are those of the unpaused sound
    plain = oracle.chain(to_oracle(rb.TestSource(x, 2, 44100).low_pass(300).amplify(0.5)))[0]
    paused = oracle.chain(to_oracle(first))[0]
    assert np.array_equal(paused[2000 + 200:2000 + 200 + 500], plain[2000:2500])

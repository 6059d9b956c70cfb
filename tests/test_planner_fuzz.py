"""Random adapter chains -- buffers and device generators, gains, filters, take_duration (with and without the fade-out filter),
delay, reverb, pause, fades, speed, explicit UniformSourceIterators, mixes of such chains two levels deep -- through the host
planner (rb_stream_plan / rb_streams_plan) against the lengths the oracle's literal pull iterators produce: the chain itself and what
two different mixers pull from it.  (A run of this fuzz found the nested take_duration padding rule.)"""
import numpy as np

import oracle
import rodio_b200 as rb
from helpers import noise, to_oracle


def random_chain(rng, sd, depth=0):
    """(source, description): a random adapter chain drawn from `rng`."""
    def base(sd):
        k=int(rng.integers(0,3)); ch=int(rng.integers(1,4)); rate=int(rng.choice([8000,22050,32000,44100,48000]))
        n=int(rng.choice([0,ch,3*ch,200*ch,1500*ch,12000*ch]))
        if k==0: return rb.SamplesBuffer(ch,rate,noise(n,sd)), f"SB{ch}/{rate}/{n}"
        if k==1: return rb.TestSource(noise(n,sd),ch,rate), f"TS{ch}/{rate}/{n}"
        f=float(rng.choice([50.0,440.0,9000.0])); m=int(rng.choice([0,1,500,5000]))
        return rb.SignalGenerator(rate,f,int(rng.integers(0,4))).take(m), f"GEN{rate}/{f}/{m}"
    if True:
        s,d=base(sd)
        for _ in range(int(rng.integers(0,4))):
            k=int(rng.integers(0,10))
            if k==0: s=s.amplify(0.7); d+=".amp"
            elif k==1: s=s.low_pass(500); d+=".lp"
            elif k==2:
                ms=int(rng.integers(1,60)); fo=bool(rng.integers(0,2)); s=s.take_duration(rb.Duration.from_millis(ms),fo); d+=f".take{ms}{'f' if fo else ''}"
            elif k==3:
                ns=int(rng.choice([12000,1000000,7000000])); s=s.delay(rb.Duration.from_nanos(ns)); d+=f".delay{ns}"
            elif k==4:
                ms=int(rng.integers(1,30)); s=s.reverb(rb.Duration.from_millis(ms),0.4); d+=f".rev{ms}"
            elif k==5:
                at=int(rng.integers(0,3000)); fr=int(rng.integers(0,50)); s=s.pause_at(at,fr); d+=f".pause{at}/{fr}"
            elif k==6: s=s.fade_in(rb.Duration.from_millis(20)); d+=".fi"
            elif k==7: s=s.speed(float(rng.choice([0.5,0.9,1.3]))); d+=".spd"
            elif k==8:
                c=int(rng.integers(1,4)); r=int(rng.choice([22050,44100,48000])); s=rb.UniformSourceIterator(s,c,r); d+=f".uni{c}/{r}"
            elif k==9 and depth<2:
                o,do=random_chain(rng,sd+1000,depth+1); s=s.mix(o); d+=f".mix({do})"
        return s,d



def test_random_chains_plan_like_the_literal_iterators():
    rng = np.random.default_rng(20260923)
    checked = refused = 0
    for t in range(500):
        s, d = random_chain(rng, 7 * t)
        want = oracle.chain(to_oracle(s))[0]
        for mixer in ((1, 48000), (2, 44100)):
            try:
                out_len, ch, rate, cl = rb.plan(s, *mixer)
            except rb.RodioB200Error:
                refused += 1
                break
            assert cl == want.size, (d, cl, want.size)
            assert out_len == oracle.chain_uniform(to_oracle(s), *mixer).size, (d, mixer)
            checked += 1
    assert checked > 900 and refused < 20, (checked, refused)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_random_chains_bit_exact_on_the_device(ctx):
    """The same random chains rendered by the CUDA path (general path, exact order) into two different mixers: bit for bit what
    the oracle's MixerSource pulls (NaNs -- a filter driven above its Nyquist rate -- compare as NaNs)."""
    from helpers import assert_bit_exact
    from rodio_b200 import capi
    rng = np.random.default_rng(777)
    done = 0
    for t in range(120):
        s, d = random_chain(rng, 50000 + 7 * t)
        mixer = ((1, 48000), (2, 44100), (3, 32000))[t % 3]
        want = oracle.mixer([to_oracle(s)], *mixer)
        try:
            with rb.Batch([s], *mixer, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
                b.upload_all()
                got = b.render_mix()
        except rb.RodioB200Error:
            continue                                     # a combination the block path refuses (documented RB_ERR_UNSUPPORTED)
        assert got.shape == want.shape, (d, got.shape, want.shape)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), d
        assert_bit_exact(np.where(nan, np.float32(0), got), np.where(nan, np.float32(0), want), d)
        done += 1
    assert done > 100

"""The committed fixtures under tests/golden/ (made by tests/golden/make_golden.py):

* reference_vectors.json -- the reference's own unit-test vectors for the path, as data -> checked against the
  CPU oracle here (CPU) and against the CUDA path through the C ABI (GPU);
* oracle_chains.json / limiter_chains.npz -- fingerprints of the oracle's output for every adapter chain of
  tests/chains.py -> pin the restatement (CPU: any drift of the oracle fails) and check the CUDA path against
  committed bytes without needing the oracle on the box (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

import rodio_b200 as rb
from chains import CHAINS, LIMIT_CHAINS
from helpers import assert_close_peak, to_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "reference_vectors.json")) as f:
    REF = json.load(f)
with open(os.path.join(GOLD, "oracle_chains.json")) as f:
    CHAIN_FP = json.load(f)["chains"]


def _f32(v):
    return np.asarray(v, dtype=np.float32)


def _fp(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return {"len": int(a.size), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
            "head_bits": [int(v) for v in a[:6].view(np.uint32)]}


# ---------------------------------------------------------------------------------------- CPU: oracle vs fixtures
def test_fixture_tables_cover_every_chain():
    assert sorted(CHAIN_FP) == sorted(CHAINS)
    with np.load(os.path.join(GOLD, "limiter_chains.npz")) as z:
        assert sorted(z.files) == sorted(LIMIT_CHAINS)


@pytest.mark.parametrize("i", range(len(REF["sample_rate_converter"])))
def test_oracle_src_reference_vectors(i):
    import oracle
    v = REF["sample_rate_converter"][i]
    out = oracle.sample_rate_converter(_f32(v["input"]), v["from"], v["to"], v["channels"])
    out = np.trunc(out) if v["trunc"] else out
    assert out.tolist() == [float(x) for x in v["output"]], v["cite"]


@pytest.mark.parametrize("i", range(len(REF["channel_count_converter"])))
def test_oracle_channel_reference_vectors(i):
    import oracle
    v = REF["channel_count_converter"][i]
    assert oracle.channel_count_converter(_f32(v["input"]), v["from"], v["to"]).tolist() == [float(x) for x in v["output"]]


@pytest.mark.parametrize("i", range(len(REF["mixer"])))
def test_oracle_mixer_reference_vectors(i):
    import oracle
    v = REF["mixer"][i]
    streams = [oracle.Stream(pcm=_f32(s["pcm"]), channels=s["channels"], sample_rate=s["rate"], effects=[],
                             span_len=len(s["pcm"])) for s in v["sources"]]     # SamplesBuffer reports its length
    out = oracle.mixer(streams, *v["mixer"])
    assert out.tolist() == [float(x) for x in v["output"]], v["cite"]


@pytest.mark.parametrize("i", range(len(REF["channel_volume"])))
def test_oracle_channel_volume_reference_vectors(i):
    import oracle
    v = REF["channel_volume"][i]
    src = rb.ChannelVolume(rb.TestSource(_f32(v["pcm"]), v["channels"], v["rate"]), v["volumes"])
    out = oracle.chain(to_oracle(src))[0]
    want = np.array([np.float32(x) for x in v["output"]], dtype=np.float32)
    assert np.allclose(out, want, rtol=0, atol=1e-6) and out.size == want.size, v["cite"]


def test_oracle_db_table():
    import oracle
    t = REF["db_table"]
    for db, lin in t["rows"]:
        r = float(oracle.db_to_linear(db)) / lin
        assert 1 - t["ratio_tolerance"] < r < 1 + t["ratio_tolerance"], (db, lin)
        if abs(db) > 1e-5:
            r = float(oracle.linear_to_db(lin)) / db
            assert 1 - t["ratio_tolerance"] < r < 1 + t["ratio_tolerance"], (db, lin)


@pytest.mark.parametrize("name", sorted(CHAINS))
def test_oracle_matches_chain_fingerprint(name):
    """The restatement has not drifted from the committed bytes."""
    import oracle
    src = CHAINS[name]()
    _, ch, rate = oracle.chain(to_oracle(src))
    want = CHAIN_FP[name]
    assert (ch, rate) == (want["channels"], want["sample_rate"])
    got = _fp(oracle.chain_uniform(to_oracle(src), ch, rate))
    assert got == {k: want[k] for k in ("len", "sha256", "head_bits")}, name


# ---------------------------------------------------------------------------------------- GPU: CUDA path vs fixtures
def _run_chain(src, ctx):
    from rodio_b200 import capi
    flags = capi.RB_KEEP_STREAM_OUTPUTS | capi.RB_NO_FUSION | capi.RB_MIX_EXACT_ORDER
    with rb.Batch([src], src.channels(), src.sample_rate(), flags=flags, ctx=ctx) as b:
        b.upload_all()
        b.render_mix_device()
        return b.read_stream(0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CHAINS))
def test_cuda_matches_chain_fingerprint(ctx, name):
    """Bit-exact against the committed fingerprint (no oracle involved on the box)."""
    want = CHAIN_FP[name]
    got = _fp(_run_chain(CHAINS[name](), ctx))
    assert got == {k: want[k] for k in ("len", "sha256", "head_bits")}, name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LIMIT_CHAINS))
def test_cuda_limiter_within_tolerance_of_fixture(ctx, name):
    with np.load(os.path.join(GOLD, "limiter_chains.npz")) as z:
        want = z[name]
    assert_close_peak(_run_chain(LIMIT_CHAINS[name](), ctx), want, 1e-5, name)   # north-star float tolerance


@pytest.mark.gpu
def test_cuda_reference_vectors(ctx):
    for v in REF["sample_rate_converter"]:
        out = rb.SampleRateConverter(_f32(v["input"]), v["from"], v["to"], v["channels"], ctx=ctx)
        out = np.trunc(out) if v["trunc"] else out
        assert out.tolist() == [float(x) for x in v["output"]], v["cite"]
    for v in REF["channel_count_converter"]:
        assert rb.ChannelCountConverter(_f32(v["input"]), v["from"], v["to"], ctx=ctx).tolist() == [float(x) for x in v["output"]]
    for v in REF["mixer"]:
        tx, rx = rb.mixer(*v["mixer"], ctx=ctx)
        for s in v["sources"]:
            tx.add(rb.SamplesBuffer(s["channels"], s["rate"], s["pcm"]))
        got = [rx.next() for _ in range(len(v["output"]) + 1)]
        assert got == [float(x) for x in v["output"]] + [None], v["cite"]
    for v in REF["channel_volume"]:
        out = _run_chain(rb.ChannelVolume(rb.TestSource(_f32(v["pcm"]), v["channels"], v["rate"]), v["volumes"]), ctx)
        assert np.allclose(out, _f32(v["output"]), rtol=0, atol=1e-6) and out.size == len(v["output"]), v["cite"]

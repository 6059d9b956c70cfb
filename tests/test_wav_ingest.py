"""WAV ingest: rb_wav_parse / rb_wav_unpack24 find the samples of a RIFF/WAVE image and name their format the
way src/decoder/wav.rs:119-151 reads them (8-bit unsigned, 16 / 24 / 32-bit integer, 32-bit float); the bytes then go to HBM as
they are and are converted on the device by dasp's rules.  Host-side tests run everywhere; the device test is marked gpu."""
import io
import struct
import wave

import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, to_oracle
from rodio_b200 import capi


def wav_bytes(width, ch, rate, frames, seed=0):
    rng = np.random.default_rng(seed)
    bio = io.BytesIO()
    w = wave.open(bio, "wb")
    w.setnchannels(ch), w.setsampwidth(width), w.setframerate(rate)
    if width == 1:
        vals = rng.integers(0, 256, frames * ch, dtype=np.uint8)
        data = vals.tobytes()
    elif width == 2:
        vals = rng.integers(-32768, 32768, frames * ch).astype("<i2")
        data = vals.tobytes()
    elif width == 3:
        vals = rng.integers(-(1 << 23), 1 << 23, frames * ch).astype(np.int32)
        data = b"".join(int(x).to_bytes(3, "little", signed=True) for x in vals)
    else:
        vals = rng.integers(-(1 << 31), 1 << 31, frames * ch).astype("<i4")
        data = vals.tobytes()
    w.writeframes(data)
    w.close()
    return bio.getvalue(), vals


def float_wav(ch, rate, samples):
    data = np.asarray(samples, "<f4").tobytes()
    fmt = struct.pack("<HHIIHH", 3, ch, rate, rate * ch * 4, ch * 4, 32)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\x00" + b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


@pytest.mark.parametrize("width,dtype,fmt", [(1, np.uint8, capi.RB_FMT_U8), (2, np.int16, capi.RB_FMT_I16),
                                             (3, np.int32, capi.RB_FMT_I24_IN_I32), (4, np.int32, capi.RB_FMT_I32)])
def test_wav_parse_finds_format_and_samples(built, width, dtype, fmt):
    image, vals = wav_bytes(width, 2, 44100, 333, seed=width)
    src = rb.wav_source(image)
    assert (src.base_channels, src.base_rate, src.pcm.dtype) == (2, 44100, np.dtype(dtype))
    assert (src.fmt_override if src.fmt_override is not None else {np.dtype(np.uint8): capi.RB_FMT_U8, np.dtype(np.int16): capi.RB_FMT_I16,
                                                                  np.dtype(np.int32): capi.RB_FMT_I32}[src.pcm.dtype]) == fmt
    assert np.array_equal(src.pcm, np.asarray(vals).astype(dtype))


def test_wav_parse_float_extra_chunks_truncation_and_errors(built):
    x = np.linspace(-1, 1, 101, dtype=np.float32)
    src = rb.wav_source(float_wav(1, 48000, x))                       # an odd-sized LIST chunk sits in front of the data
    assert src.pcm.dtype == np.float32 and np.array_equal(src.pcm, x) and src.base_rate == 48000
    image, vals = wav_bytes(2, 2, 22050, 100)
    cut = rb.wav_source(image[:-7])                                   # a truncated file: whole frames only
    assert cut.pcm.size == 196 and np.array_equal(cut.pcm, vals[:196])
    for bad in (b"RIFF\x00\x00\x00\x00WAVX", image[:20], b"junk"):
        with pytest.raises(rb.RodioB200Error):
            rb.wav_source(bad)
    adpcm = bytearray(image)
    adpcm[20:22] = struct.pack("<H", 2)                               # a compressed format tag: rodio's decoder does not read it either
    with pytest.raises(rb.RodioB200Error) as e:
        rb.wav_source(bytes(adpcm))
    assert e.value.status == capi.RB_ERR_UNSUPPORTED


@pytest.mark.gpu
@pytest.mark.parametrize("width", [1, 2, 3, 4])
def test_wav_to_hbm_to_mix(ctx, width):
    """assets/*.wav shape: a stereo 44.1 kHz file goes to HBM in its own format, is converted on the device and resampled into a
    48 kHz stereo mixer beside a second file -- bit for bit the oracle's conversion + chain."""
    srcs = [rb.UniformSourceIterator(rb.wav_source(wav_bytes(width, 2, 44100, 2000 + 100 * i, seed=10 * width + i)[0]), 2, 48000).amplify(0.5) for i in range(3)]
    want = oracle.mixer([to_oracle(s) for s in srcs], 2, 48000)
    with rb.Batch(srcs, 2, 48000, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
    assert_bit_exact(got, want, f"{8 * width}-bit WAV through the mixer")

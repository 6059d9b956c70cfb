/* Plain-C use of the streaming boundary (include/rodio_b200.h): a stereo 48 kHz mixer fed by one 44.1 kHz stereo source and
 * one 48 kHz mono source that arrive in 10 ms blocks -- the block form of
 *     let (mixer, mut out) = rodio::mixer::mixer(nz!(2), nz!(48_000));
 *     mixer.add(music.low_pass(200).amplify(0.8));   mixer.add(voice);          // src/mixer.rs:25-66
 *     while let Some(sample) = out.next() { .. }                                 // src/mixer.rs:120-136
 * Build:  gcc -std=c11 -Iinclude examples/stream_mixer.c -Lrodio_b200 -l:librodio_b200.so -Wl,-rpath,$PWD/rodio_b200 -lm -o stream_mixer
 * (needs a CUDA device at run time: rb_context_create fails loudly without one). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rodio_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        rb_status s_ = (call);                                                               \
        if (s_ != RB_OK) {                                                                   \
            fprintf(stderr, "%s: %s (%s)\n", #call, rb_status_string(s_), rb_last_error()); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(int argc, char** argv) {
    const int blocks = argc > 1 ? atoi(argv[1]) : 100;   /* 10 ms each: one second by default */
    rb_context* ctx = NULL;
    CHECK(rb_context_create(0, &ctx));

    /* the two chains: what the mixer's UniformSourceIterator and the user's adapters do to each source */
    rb_effect music_fx[3], voice_fx[1];
    memset(music_fx, 0, sizeof music_fx), memset(voice_fx, 0, sizeof voice_fx);
    music_fx[0].kind = RB_FX_UNIFORM, music_fx[0].u32[0] = 2, music_fx[0].u32[1] = 48000;
    music_fx[1].kind = RB_FX_LOW_PASS, music_fx[1].u32[0] = 200, music_fx[1].f32[0] = 0.5f;   /* low_pass(200): q = 0.5 */
    music_fx[2].kind = RB_FX_AMPLIFY, music_fx[2].f32[0] = 0.8f;
    voice_fx[0].kind = RB_FX_UNIFORM, voice_fx[0].u32[0] = 2, voice_fx[0].u32[1] = 48000;
    rb_stream_desc descs[2];
    memset(descs, 0, sizeof descs);
    descs[0].sample_rate = 44100, descs[0].channels = 2, descs[0].format = RB_FMT_F32, descs[0].n_effects = 3, descs[0].effects = music_fx;
    descs[1].sample_rate = 48000, descs[1].channels = 1, descs[1].format = RB_FMT_F32, descs[1].n_effects = 1, descs[1].effects = voice_fx;

    rb_session* mixer = NULL;
    CHECK(rb_session_create(ctx, 2, 48000, descs, 2, /* fifo_frames */ 4096, /* max_block_frames */ 480, &mixer));

    enum { MUSIC_BLOCK = 441, VOICE_BLOCK = 480 };
    float music[MUSIC_BLOCK * 2], voice[VOICE_BLOCK], out[480 * 2];
    double peak = 0.0;
    uint64_t total = 0;
    int ended = 0;
    for (int b = 0; !ended; b++) {
        if (b < blocks) {   /* "decode" the next 10 ms of both sources */
            for (int i = 0; i < MUSIC_BLOCK; i++) {
                const double t = (b * MUSIC_BLOCK + i) / 44100.0;
                music[2 * i] = (float)(0.5 * sin(2 * M_PI * 110.0 * t)), music[2 * i + 1] = (float)(0.5 * sin(2 * M_PI * 165.0 * t));
            }
            for (int i = 0; i < VOICE_BLOCK; i++) voice[i] = (float)(0.25 * sin(2 * M_PI * 440.0 * (b * VOICE_BLOCK + i) / 48000.0));
            CHECK(rb_session_push(mixer, 0, music, MUSIC_BLOCK, b == blocks - 1));
            CHECK(rb_session_push(mixer, 1, voice, VOICE_BLOCK, b == blocks - 1));
        }
        for (;;) {   /* hand the mixer output on (to a device callback, a file ...) as it becomes available */
            uint64_t n = 0;
            CHECK(rb_session_render(mixer, out, 480, &n, &ended));
            for (uint64_t i = 0; i < 2 * n; i++) peak = fabs(out[i]) > peak ? fabs(out[i]) : peak;
            total += n;
            if (n == 0 || ended) break;
        }
    }
    printf("%llu stereo frames at 48 kHz, peak %.3f\n", (unsigned long long)total, peak);
    rb_session_destroy(mixer);
    rb_context_destroy(ctx);
    return 0;
}

// Entry points of the host adapter (demod_adapter.cpp) and the binding layer that lets ONE source compile against either
//   * the reference's own structs: -DABG_WITH_REFERENCE_HEADERS, include path = reference src/ with
//     integration/reference_b200.patch applied (that is how it is built inside the reference tree, and how
//     `make refcheck` / tests/test_reference_binding.py prove the field names, types and offsets it relies on), or
//   * the mirror in airband_host.h (default: the reference's third-party headers are not installed here, so the tests of
//     this repository build the adapter against a restatement of the same fields).
#pragma once
#include <stdio.h>
#ifdef ABG_WITH_REFERENCE_HEADERS
#include "rtl_airband.h"
#include "airband_b200.h"
extern int devices_running;            // rtl_airband.cpp:74 (`static` removed by the patch)
extern "C" int b200_fm_demod(void);    // rtl_airband.cpp:88-89 `fm_demod` as an int (added by the patch)
struct b200_globals {                  // adapter-only settings; everything else is the reference's own globals
    volatile int engine_ready;
    int wait_for_consumer;
    int max_batches_per_run;
    char last_error[512];
};
extern b200_globals g_b200;
namespace bd {
inline device_t* devs() { return devices; }
inline size_t fft() { return fft_size; }
inline int wave_rate() { return WAVE_RATE; }
inline int fm_demod_algo() { return b200_fm_demod(); }
inline volatile int& exit_flag() { return do_exit; }
inline int& running() { return devices_running; }
inline mixer_t* mixer_array() { return mixers; }
inline int mixer_n() { return mixer_count; }
inline float channel_alpha(const channel_t* ch) {
#ifdef NFM
    return ch->alpha;
#else
    (void)ch;
    return 0.0f;
#endif
}
inline int is_nfm(const freq_t* f) {
#ifdef NFM
    return f->modulation == MOD_NFM;
#else
    (void)f;
    return 0;
#endif
}
inline void device_failed(device_t* dev) { disable_device_outputs(dev); }  // rtl_airband.cpp:386
inline void wait_signal_ms(Signal* s, int) { s->wait(); }
}  // namespace bd
#else
#include "airband_host.h"
namespace bd {
inline device_t* devs() { return g_b200.devices; }
inline size_t fft() { return g_b200.fft_size; }
inline int wave_rate() { return g_b200.wave_rate; }
inline int fm_demod_algo() { return g_b200.fm_demod; }
inline volatile int& exit_flag() { return g_b200.do_exit; }
inline volatile int& running() { return g_b200.devices_running; }
inline mixer_t* mixer_array() { return g_b200.mixers; }
inline int mixer_n() { return g_b200.mixer_count; }
inline float channel_alpha(const channel_t* ch) { return ch->alpha; }
inline int is_nfm(const freq_t* f) { return f->modulation == MOD_NFM; }
inline void device_failed(device_t* dev) {
    if (g_b200.on_device_failed) g_b200.on_device_failed(dev);
}
}  // namespace bd
#endif

// Drop-in for `void* demodulate(void* params)` (reference src/rtl_airband.cpp:286, started at :1111).
extern "C" ABG_API void* demodulate_b200(void* params);
// 1 when the mixer is summed on the GPU by the running demod thread (process_outputs() then skips mixer_put_samples for its inputs).
extern "C" ABG_API int b200_mixer_is_gpu(const mixer_t* m);
// Squelch read-outs of one device's channels for the stats file / TUI (output.cpp:598-869, rtl_airband.cpp:632-643).
extern "C" ABG_API int b200_refresh_stats(abg_engine* eng, int dev_local, device_t* dev);
// What process_outputs() writes for an O_RAWFILE output (output.cpp:519-522): one batch of interleaved float32 I/Q (.cf32).
extern "C" ABG_API size_t b200_write_rawfile(FILE* f, const channel_t* channel, int wave_batch);

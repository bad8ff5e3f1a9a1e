/*
 * airband_b200_host.h — the two small structs the WITH_B200 build adds to the reference's freq_t
 * (reference src/rtl_airband.h:223-233; integration/reference_b200.patch), shared by that patch and by this repository's
 * mirror of the reference structs (rtlsdr-airband_b200/host/airband_host.h).
 *
 * Why they exist: Squelch / NotchFilter / LowpassFilter keep what config.cpp gave them in private members without getters
 * (reference src/squelch.h:117-158, src/filters.h:32-38,55-60), so the values parse_channels() passes to them
 * (reference src/config.cpp:437-619) are recorded next to the objects; and the Squelch read-outs the stats file / TUI print
 * (reference src/output.cpp:598-869, src/rtl_airband.cpp:632-643) come back from the engine instead of the CPU objects.
 */
#ifndef AIRBAND_B200_HOST_H
#define AIRBAND_B200_HOST_H
#include <stddef.h>

struct b200_freq_cfg {
    float squelch_level;  /* > 0: set_squelch_level_threshold(level) was called with it (config.cpp:447-466) */
    float squelch_snr_db; /* >= 0: set_squelch_snr_threshold(db) was called (config.cpp:494,509); mk_freqlist() starts it at -1 */
    float notch_hz, notch_q; /* NotchFilter(freq, WAVE_RATE, q) (config.cpp:541,557) */
    float ctcss_hz;       /* set_ctcss_freq(freq, WAVE_RATE) (config.cpp:575,584) */
    float lowpass_hz;     /* LowpassFilter(bandwidth / 2, WAVE_RATE) (config.cpp:604,615) */
};
struct b200_freq_stats {
    float noise_level, signal_level, squelch_level;                /* Squelch getters, squelch.h:89-91 */
    float noise_level_dbfs, signal_level_dbfs, squelch_level_dbfs; /* level_to_dBFS(), util.cpp:178-180 */
    size_t open_count, flappy_count, ctcss_count, no_ctcss_count;  /* squelch.h:93-96 */
};
#endif

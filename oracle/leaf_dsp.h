// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// CPU restatement of the reference's four leaf DSP classes used inside demodulate():
//   Squelch        <- reference src/squelch.h:69-179, src/squelch.cpp:36-518
//   CTCSS & co     <- reference src/ctcss.h:26-96,   src/ctcss.cpp:31-172
//   NotchFilter    <- reference src/filters.h:25-40, src/filters.cpp:30-64
//   LowpassFilter  <- reference src/filters.h:42-63, src/filters.cpp:67-163
//
// The public method names match the reference so that oracle/demod_loop.inc compiles unchanged against
// either these restatements (libairband_oracle.so) or the reference's own classes compiled in place from
// /root/reference/src (oracle/_ref/libairband_ref.so).  tests/test_oracle_leaf_vs_ref.py drives both
// through the per-sample harness and requires bit-identical traces — that is what pins this file.
//
// Everything lives in namespace abo so the two libraries can be loaded into one process.
#pragma once
#include <complex>
#include <cstddef>
#include <vector>

namespace abo {

// ---- Goertzel tone detector (ctcss.cpp:31-60) -------------------------------------------------------------
struct ToneDetector {
    float tone_freq, magnitude, coeff;
    int window, count;
    float q0, q1, q2;
    ToneDetector(float tone_freq_hz, float sample_rate, int window_size);
    void process_sample(const float& sample);
    void reset();
};

// ---- CTCSS decision over a bank of detectors (ctcss.cpp:62-172) ------------------------------------------
class CTCSS {
   public:
    CTCSS() : enabled_(false), found_count_(0), not_found_count_(0) {}
    CTCSS(const float& ctcss_freq, const float& sample_rate, int window_size);
    void process_audio_sample(const float& sample);
    void reset();
    const size_t& found_count() const { return found_count_; }
    const size_t& not_found_count() const { return not_found_count_; }
    bool is_enabled() const { return enabled_; }
    bool enough_samples() const { return enough_samples_; }
    bool has_tone() const { return !enabled_ || has_tone_; }
    size_t tone_count() const { return bank_.size(); }
    static const float standard_tones[51];

   private:
    bool enabled_;
    float ctcss_freq_ = 0;
    int window_size_ = 0;
    size_t found_count_, not_found_count_;
    std::vector<ToneDetector> bank_;  // bank_[0] is always the wanted tone (ctcss.cpp:103)
    bool enough_samples_ = false;
    int sample_count_ = 0;
    bool has_tone_ = false;
};

// ---- Squelch state machine (squelch.cpp) -------------------------------------------------------------------
class Squelch {
   public:
    Squelch();
    void set_squelch_level_threshold(const float& level);
    void set_squelch_snr_threshold(const float& db);
    void set_ctcss_freq(const float& ctcss_freq, const float& sample_rate);

    void process_raw_sample(const float& sample);
    void process_filtered_sample(const float& sample);
    void process_audio_sample(const float& sample);

    bool is_open() const;
    bool should_filter_sample();
    bool should_process_audio();
    bool first_open_sample() const;
    bool last_open_sample() const;
    bool signal_outside_filter();

    const float& noise_level() const { return noise_floor_; }
    const float& signal_level() const { return pre_.full; }
    const float& squelch_level();
    const size_t& open_count() const { return open_count_; }
    const size_t& flappy_count() const { return flappy_count_; }
    const size_t& ctcss_count() const { return ctcss_slow_.found_count(); }
    const size_t& no_ctcss_count() const { return ctcss_slow_.not_found_count(); }

    // oracle-only introspection for traces (not in the reference API)
    int debug_state() const { return (int)cur_; }
    int debug_next_state() const { return (int)next_; }
    float debug_pre_capped() const { return pre_.capped; }
    float debug_post_capped() const { return post_.capped; }

   private:
    enum State { CLOSED, OPENING, CLOSING, LOW_SIGNAL_ABORT, OPEN };  // squelch.h:104-110
    struct Avg {
        float full, capped;
    };
    float noise_floor_;
    bool manual_;
    float manual_level_;
    float normal_ratio_, flappy_ratio_;
    float avg_cap_;
    Avg pre_, post_;
    float level_cache_;
    bool using_post_;
    float pre_vs_post_;
    int open_delay_, close_delay_, low_signal_abort_;
    State next_, cur_;
    int delay_;
    size_t open_count_, sample_count_, flappy_count_;
    int low_signal_count_;
    size_t recent_sample_size_, flap_opens_threshold_, recent_open_count_, closed_sample_count_;
    int buf_size_, head_, tail_;
    std::vector<float> buf_;
    CTCSS ctcss_fast_, ctcss_slow_;

    void set_state(State s);
    void update_current_state();
    bool has_pre_signal();
    bool has_post_signal();
    bool has_signal();
    void calc_noise_floor();
    void calc_avg_cap();
    void update_avg(Avg& a, const float& sample);
    bool flapping() const { return recent_open_count_ >= flap_opens_threshold_; }
};

// ---- audio notch (filters.cpp:30-64) -----------------------------------------------------------------------
class NotchFilter {
   public:
    NotchFilter() : enabled_(false) {}
    NotchFilter(float notch_freq, float sample_freq, float q);
    void apply(float& value);
    bool enabled() { return enabled_; }

   private:
    bool enabled_;
    float e = 0, p = 0, d[3] = {0, 0, 0}, x[3] = {0, 0, 0}, y[3] = {0, 0, 0};
};

// ---- complex 2-pole Bessel low-pass (filters.cpp:67-163) ---------------------------------------------------
class LowpassFilter {
   public:
    LowpassFilter() : enabled_(false) {}
    LowpassFilter(float freq, float sample_freq);
    void apply(float& r, float& j);
    bool enabled() const { return enabled_; }
    // oracle-only: coefficients for cross-checks against the CUDA host-side design code
    float debug_gain() const { return gain; }
    float debug_yc(int i) const { return ycoeffs[i]; }

   private:
    bool enabled_;
    float ycoeffs[3] = {0, 0, 0};
    float gain = 1;
    std::complex<float> xv[3], yv[3];
};

}  // namespace abo

// ORACLE — TEST INFRASTRUCTURE ONLY (see leaf_dsp.h).  Restated leaf DSP of the reference demod path.
// Every arithmetic expression keeps the reference's operand types, association and rounding points
// (float vs double), because squelch decisions are hard compares on these values.
#include "leaf_dsp.h"

#include <math.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace abo {

// =============================================================================================================
// ToneDetector — ctcss.cpp:31-60
// =============================================================================================================
ToneDetector::ToneDetector(float tone_freq_hz, float sample_rate, int window_size) {
    tone_freq = tone_freq_hz;
    magnitude = 0.0;
    window = window_size;
    // ctcss.cpp:37-39: int*float/float stays float, +0.5 promotes to double, truncation to int;
    // omega is rounded to float before the double cos().
    int k = (0.5 + window_size * tone_freq_hz / sample_rate);
    float omega = (2.0 * M_PI * k) / window_size;
    coeff = 2.0 * cos(omega);
    reset();
}

void ToneDetector::process_sample(const float& sample) {
    q0 = coeff * q1 - q2 + sample;  // ctcss.cpp:45
    q2 = q1;
    q1 = q0;
    count++;
    if (count == window) {
        magnitude = q1 * q1 + q2 * q2 - q1 * q2 * coeff;  // ctcss.cpp:51
        count = 0;
    }
}

void ToneDetector::reset() {
    count = 0;
    q0 = q1 = q2 = 0.0;
}

// =============================================================================================================
// CTCSS — ctcss.cpp:92-172
// =============================================================================================================
const float CTCSS::standard_tones[51] = {67.0,  69.3,  71.9,  74.4,  77.0,  79.7,  82.5,  85.4,  88.5,  91.5,  94.8,  97.4,  100.0,
                                         103.5, 107.2, 110.9, 114.8, 118.8, 123.0, 127.3, 131.8, 136.5, 141.3, 146.2, 150.0, 151.4,
                                         156.7, 159.8, 162.2, 165.5, 167.9, 171.3, 173.8, 177.3, 179.9, 183.5, 186.2, 189.9, 192.8,
                                         196.6, 199.5, 203.5, 206.5, 210.7, 218.1, 225.7, 229.1, 233.6, 241.8, 250.3, 254.1};

CTCSS::CTCSS(const float& ctcss_freq, const float& sample_rate, int window_size)
    : enabled_(true), ctcss_freq_(ctcss_freq), window_size_(window_size), found_count_(0), not_found_count_(0) {
    // wanted tone first, then every standard tone not within 5 Hz of it; a candidate whose Goertzel
    // coefficient collides bit-for-bit with one already in the bank is dropped (ctcss.cpp:62-73,98-111)
    auto try_add = [&](float f) {
        ToneDetector cand(f, sample_rate, window_size_);
        for (const ToneDetector& t : bank_)
            if (cand.coeff == t.coeff) return;
        bank_.push_back(cand);
    };
    try_add(ctcss_freq);
    for (float tone : standard_tones) {
        if (std::abs(ctcss_freq - tone) < 5) continue;
        try_add(tone);
    }
    reset();
}

void CTCSS::process_audio_sample(const float& sample) {
    if (!enabled_) return;
    for (ToneDetector& t : bank_) t.process_sample(sample);
    sample_count_++;
    if (sample_count_ < window_size_) return;
    enough_samples_ = true;

    // ctcss.cpp:78-90,132-156: sequential float sum in bank order, mean = sum / count; the wanted tone must
    // equal the maximum power and exceed the mean.  (The reference sorts and takes element 0; the max is the
    // same value.  It looks the wanted tone up by frequency; the first match in the sorted list with that
    // frequency is bank_[0] unless a *different* detector has an identical frequency, which the 5 Hz rule
    // excludes.)
    float total = 0.0;
    float maxp = bank_[0].magnitude;
    for (const ToneDetector& t : bank_) {
        total += t.magnitude;
        if (t.magnitude > maxp) maxp = t.magnitude;
    }
    float avg = total / bank_.size();
    float want = bank_[0].magnitude;
    if (want == maxp && want > avg) {
        has_tone_ = true;
        found_count_++;
    } else {
        has_tone_ = false;
        not_found_count_++;
    }
    for (ToneDetector& t : bank_) t.reset();
    sample_count_ = 0;
}

void CTCSS::reset() {
    if (!enabled_) return;
    for (ToneDetector& t : bank_) t.reset();
    enough_samples_ = false;
    sample_count_ = 0;
    has_tone_ = false;
}

// =============================================================================================================
// Squelch — squelch.cpp
// =============================================================================================================
Squelch::Squelch() {  // squelch.cpp:36-82
    noise_floor_ = 5.0f;
    manual_level_ = 0.0f;  // reference leaves it unset until after the next call; never read while !manual_
    set_squelch_snr_threshold(9.54f);
    manual_level_ = -1.0;
    pre_ = {0.001f, 0.001f};
    post_ = {0.001f, 0.001f};
    level_cache_ = 0.0f;
    using_post_ = false;
    pre_vs_post_ = 0.9f;
    open_delay_ = 197;
    close_delay_ = 197;
    low_signal_abort_ = 88;
    next_ = CLOSED;
    cur_ = CLOSED;
    delay_ = 0;
    open_count_ = 0;
    sample_count_ = (size_t)-1;
    flappy_count_ = 0;
    low_signal_count_ = 0;
    recent_sample_size_ = 1000;
    flap_opens_threshold_ = 3;
    recent_open_count_ = 0;
    closed_sample_count_ = 0;
    buf_size_ = 102;
    head_ = 0;
    tail_ = 1;
    buf_.assign(buf_size_, 0.0f);
}

void Squelch::set_squelch_level_threshold(const float& level) {  // squelch.cpp:84-96
    if (level > 0) {
        manual_ = true;
        manual_level_ = level;
    } else {
        manual_ = false;
    }
    calc_avg_cap();
}

void Squelch::set_squelch_snr_threshold(const float& db) {  // squelch.cpp:98-108
    manual_ = false;
    normal_ratio_ = pow(10.0, db / 20.0);
    flappy_ratio_ = normal_ratio_ * 0.9f;
    calc_avg_cap();
}

void Squelch::set_ctcss_freq(const float& ctcss_freq, const float& sample_rate) {  // squelch.cpp:110-116
    ctcss_fast_ = CTCSS(ctcss_freq, sample_rate, sample_rate * 0.05);
    ctcss_slow_ = CTCSS(ctcss_freq, sample_rate, sample_rate * 0.4);
}

bool Squelch::is_open() const {  // squelch.cpp:118-134
    if (cur_ == OPEN || cur_ == CLOSING) {
        if (ctcss_slow_.is_enabled()) {
            if (ctcss_slow_.enough_samples()) return ctcss_slow_.has_tone();
            return ctcss_fast_.has_tone();
        }
        return true;
    }
    return false;
}

bool Squelch::should_filter_sample() { return ((has_pre_signal() || cur_ != CLOSED) && cur_ != LOW_SIGNAL_ABORT); }
bool Squelch::should_process_audio() { return (cur_ == OPEN || cur_ == CLOSING); }
bool Squelch::first_open_sample() const { return (cur_ != OPEN && next_ == OPEN); }
bool Squelch::last_open_sample() const {
    return (cur_ == CLOSING && next_ == CLOSED) || (cur_ != LOW_SIGNAL_ABORT && next_ == LOW_SIGNAL_ABORT);
}
bool Squelch::signal_outside_filter() { return (using_post_ && has_pre_signal() && !has_post_signal()); }

const float& Squelch::squelch_level() {  // squelch.cpp:164-177
    if (manual_) return manual_level_;
    if (level_cache_ == 0.0f) {
        if (flapping() && flappy_ratio_ < normal_ratio_)
            level_cache_ = flappy_ratio_ * noise_floor_;
        else
            level_cache_ = normal_ratio_ * noise_floor_;
    }
    return level_cache_;
}

void Squelch::process_raw_sample(const float& sample) {  // squelch.cpp:195-246
    update_current_state();
    sample_count_++;
    if (sample_count_ % 16 == 0) calc_noise_floor();
    update_avg(pre_, sample);
    buf_[head_] = pre_.capped * pre_vs_post_;

    if (cur_ == OPEN && !has_signal()) set_state(CLOSING);
    if (cur_ == CLOSED && has_signal()) set_state(OPENING);

    if (cur_ != CLOSED && cur_ != LOW_SIGNAL_ABORT) {
        if (sample >= squelch_level()) {
            low_signal_count_ = 0;
        } else {
            low_signal_count_++;
            if (low_signal_count_ >= low_signal_abort_) set_state(LOW_SIGNAL_ABORT);
        }
    }
}

void Squelch::process_filtered_sample(const float& sample) {  // squelch.cpp:248-276
    if (!should_filter_sample()) return;
    if (cur_ == OPENING) {
        if (delay_ < buf_size_) return;
        if (delay_ == buf_size_) post_ = {buf_[tail_], buf_[tail_]};
    }
    using_post_ = true;
    update_avg(post_, sample);
    if (post_.capped < buf_[tail_]) set_state(CLOSED);
}

void Squelch::process_audio_sample(const float& sample) {  // squelch.cpp:278-295
    if (!ctcss_slow_.is_enabled()) return;
    if (cur_ != CLOSED) {
        ctcss_slow_.process_audio_sample(sample);
        if (!ctcss_slow_.enough_samples()) ctcss_fast_.process_audio_sample(sample);
    }
}

void Squelch::set_state(State s) {  // squelch.cpp:297-361 (illegal transitions are redirected)
    if (cur_ == CLOSED && s == CLOSING)
        s = CLOSED;
    else if (cur_ == CLOSED && s == LOW_SIGNAL_ABORT)
        s = CLOSED;
    else if (cur_ == CLOSED && s == OPEN)
        s = OPENING;
    else if (cur_ == OPENING && s == LOW_SIGNAL_ABORT)
        s = CLOSED;
    else if (cur_ == LOW_SIGNAL_ABORT && s != LOW_SIGNAL_ABORT && s != CLOSED)
        s = CLOSED;
    else if (cur_ == OPEN && s == CLOSED)
        s = CLOSING;
    else if (cur_ == OPEN && s == OPENING)
        s = OPEN;
    next_ = s;
}

void Squelch::update_current_state() {  // squelch.cpp:363-460
    if (next_ == OPENING) {
        if (cur_ != OPENING) {
            delay_ = 0;
            low_signal_count_ = 0;
            using_post_ = false;
            cur_ = next_;
        } else {
            delay_++;
            if (delay_ >= open_delay_) {
                if (closed_sample_count_ < recent_sample_size_) {
                    recent_open_count_++;
                    if (flapping()) flappy_count_++;
                    level_cache_ = 0.0f;
                }
                next_ = has_signal() ? OPEN : CLOSED;
            }
        }
    } else if (next_ == CLOSING) {
        if (cur_ != CLOSING) {
            delay_ = 0;
            cur_ = next_;
        } else {
            delay_++;
            if (delay_ >= close_delay_) {
                if (!has_signal()) {
                    next_ = CLOSED;
                } else {
                    cur_ = OPEN;
                    next_ = OPEN;
                }
            }
        }
    } else if (next_ == LOW_SIGNAL_ABORT) {
        if (cur_ != LOW_SIGNAL_ABORT) {
            if (cur_ != CLOSING) delay_ = 0;
            cur_ = next_;
        } else {
            delay_++;
            if (delay_ >= close_delay_) next_ = CLOSED;
        }
    } else if (next_ == OPEN && cur_ != OPEN) {
        open_count_++;
        cur_ = next_;
    } else if (next_ == CLOSED && cur_ != CLOSED) {
        using_post_ = false;
        closed_sample_count_ = 0;
        cur_ = next_;
        ctcss_fast_.reset();
        ctcss_slow_.reset();
    } else if (next_ == CLOSED && cur_ == CLOSED) {
        if (closed_sample_count_ < recent_sample_size_) {
            closed_sample_count_++;
        } else if (closed_sample_count_ == recent_sample_size_) {
            recent_open_count_ = 0;
            level_cache_ = 0.0f;
        }
    } else {
        cur_ = next_;
    }
    tail_ = (tail_ + 1) % buf_size_;
    head_ = (head_ + 1) % buf_size_;
}

bool Squelch::has_pre_signal() { return pre_.capped >= squelch_level(); }
bool Squelch::has_post_signal() { return using_post_ && post_.capped >= buf_[tail_]; }
bool Squelch::has_signal() {
    if (using_post_) return has_pre_signal() && has_post_signal();
    return has_pre_signal();
}

void Squelch::calc_noise_floor() {  // squelch.cpp:477-490
    static const float decay_factor = 0.97f;
    static const float new_factor = 1.0 - decay_factor;
    noise_floor_ = noise_floor_ * decay_factor + std::min(pre_.capped, noise_floor_) * new_factor + 1e-6f;
    calc_avg_cap();
    level_cache_ = 0.0f;
}

void Squelch::calc_avg_cap() {  // squelch.cpp:492-499
    if (manual_)
        avg_cap_ = 1.5f * manual_level_;
    else
        avg_cap_ = 1.5f * normal_ratio_ * noise_floor_;
}

void Squelch::update_avg(Avg& a, const float& sample) {  // squelch.cpp:501-514
    static const float decay_factor = 0.99f;
    static const float new_factor = 1.0 - decay_factor;
    a.full = a.full * decay_factor + sample * new_factor;
    if (a.capped >= avg_cap_ && sample >= avg_cap_)
        a.capped = avg_cap_;
    else
        a.capped = std::min(avg_cap_, a.capped * decay_factor + sample * new_factor);
}

// =============================================================================================================
// NotchFilter — filters.cpp:30-64
// =============================================================================================================
NotchFilter::NotchFilter(float notch_freq, float sample_freq, float q) : enabled_(true) {
    if (notch_freq <= 0.0) {
        enabled_ = false;
        return;
    }
    float wo = 2 * M_PI * (notch_freq / sample_freq);  // float ratio, double product, rounded to float
    e = 1 / (1 + tan(wo / (q * 2)));                    // float arg promoted to double tan, result to float
    p = cos(wo);
    d[0] = e;
    d[1] = 2 * e * p;
    d[2] = (2 * e - 1);
}

void NotchFilter::apply(float& value) {
    if (!enabled_) return;
    x[0] = x[1];
    x[1] = x[2];
    x[2] = value;
    y[0] = y[1];
    y[1] = y[2];
    y[2] = d[0] * x[2] - d[1] * x[1] + d[0] * x[0] + d[1] * y[1] - d[2] * y[0];
    value = y[2];
}

// =============================================================================================================
// LowpassFilter — filters.cpp:67-163 (mkfilter-style 2-pole Bessel via bilinear transform, designed in double)
// =============================================================================================================
namespace {
typedef std::complex<double> cd;
cd blt(cd pz) { return (2.0 + pz) / (2.0 - pz); }
void multin(cd w, int npz, cd coeffs[]) {
    cd nw = -w;
    for (int i = npz; i >= 1; i--) coeffs[i] = (nw * coeffs[i]) + coeffs[i - 1];
    coeffs[0] = nw * coeffs[0];
}
void expand(cd pz[], int npz, cd coeffs[]) {
    coeffs[0] = 1.0;
    for (int i = 0; i < npz; i++) coeffs[i + 1] = 0.0;
    for (int i = 0; i < npz; i++) multin(pz[i], npz, coeffs);
    for (int i = 0; i < npz + 1; i++) {
        if (fabs(coeffs[i].imag()) > 1e-10) {
            fprintf(stderr, "oracle: lowpass design: coeff of z^%d is not real\n", i);
            abort();
        }
    }
}
cd eval(cd coeffs[], int npz, cd z) {
    cd sum(0.0);
    for (int i = npz; i >= 0; i--) sum = (sum * z) + coeffs[i];
    return sum;
}
}  // namespace

LowpassFilter::LowpassFilter(float freq, float sample_freq) : enabled_(true) {
    if (freq <= 0.0) {
        enabled_ = false;
        return;
    }
    double raw_alpha = (double)freq / sample_freq;
    double warped_alpha = tan(M_PI * raw_alpha) / M_PI;
    cd zeros[2] = {-1.0, -1.0};
    cd poles[2];
    poles[0] = blt(M_PI * 2 * warped_alpha * cd(-1.10160133059e+00, 6.36009824757e-01));
    poles[1] = blt(M_PI * 2 * warped_alpha * conj(cd(-1.10160133059e+00, 6.36009824757e-01)));
    cd top[3], bot[3];
    expand(zeros, 2, top);
    expand(poles, 2, bot);
    cd g = eval(top, 2, 1.0) / eval(bot, 2, 1.0);
    gain = hypot(g.imag(), g.real());
    for (int i = 0; i <= 2; i++) ycoeffs[i] = -(bot[i].real() / bot[2].real());
}

void LowpassFilter::apply(float& r, float& j) {
    if (!enabled_) return;
    std::complex<float> input(r, j);
    xv[0] = xv[1];
    xv[1] = xv[2];
    xv[2] = input / gain;
    yv[0] = yv[1];
    yv[1] = yv[2];
    yv[2] = (xv[0] + xv[2]) + (2.0f * xv[1]) + (ycoeffs[0] * yv[0]) + (ycoeffs[1] * yv[1]);
    r = yv[2].real();
    j = yv[2].imag();
}

}  // namespace abo

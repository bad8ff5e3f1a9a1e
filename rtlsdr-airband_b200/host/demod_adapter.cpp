// demodulate_b200(): the reference's demod thread function re-expressed over the C ABI of the B200 engine.
// Same contract as demodulate() (reference src/rtl_airband.cpp:286-672): it owns devices[device_start..device_end),
// consumes each input ring under the reference's locking discipline (:370-375, bufs advanced without the lock, :669),
// follows the input state machine (:377-391), delivers finished batches into channel_t.waveout / iq_out /
// axcindicate, bumps active_counter (:645-647), raises waveavail or counts an overrun (:649-654), and signals the
// output thread (:662).  Fatal engine errors are reported like the VideoCore branch does (:296-310): message + exit.
#include "airband_host.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <vector>

b200_globals g_b200;

void Signal::wait_ms(int ms) {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_nsec += (long)(ms % 1000) * 1000000L;
    ts.tv_sec += ms / 1000 + ts.tv_nsec / 1000000000L;
    ts.tv_nsec %= 1000000000L;
    pthread_mutex_lock(&mutex_);
    pthread_cond_timedwait(&cond_, &mutex_, &ts);
    pthread_mutex_unlock(&mutex_);
}

// reference src/input-helpers.cpp:37-63
void circbuffer_append(input_t* const input, unsigned char* buf, size_t len) {
    if (len == 0) return;
    pthread_mutex_lock(&input->buffer_lock);
    const size_t tail = 2 * input->bytes_per_sample * g_b200.fft_size;
    size_t space_left = input->buf_size - input->bufe;
    if (space_left >= len) {
        memcpy(input->buffer + input->bufe, buf, len);
        if (input->bufe == 0) memcpy(input->buffer + input->buf_size, input->buffer, std::min(len, tail));
    } else {
        memcpy(input->buffer + input->bufe, buf, space_left);
        memcpy(input->buffer, buf + space_left, len - space_left);
        memcpy(input->buffer + input->buf_size, input->buffer, std::min(len - space_left, tail));
    }
    size_t old_end = input->bufe;
    input->bufe = (input->bufe + len) % input->buf_size;
    if (old_end < input->bufs && input->bufe >= input->bufs) input->overflow_count++;
    pthread_mutex_unlock(&input->buffer_lock);
}

static void fatal(const char* what) {
    snprintf(g_b200.last_error, sizeof(g_b200.last_error), "%s: %s", what, abg_last_error());
    fprintf(stderr, "%s\n", g_b200.last_error);  // log(LOG_CRIT, ...) in the reference tree
    g_b200.do_exit = 1;                          // error() = _Exit(1) there; the test harness wants to survive
}

// channel_t / freq_t -> abg_channel_cfg (what INTEGRATION.md calls b200_channel_cfg)
static void fill_channel_cfg(const device_t* dev, int c, const freq_t* f, abg_channel_cfg* out) {
    const channel_t* ch = dev->channels + c;
    abg_channel_cfg& o = *out;
    memset(&o, 0, sizeof(o));
    o.bin = (int32_t)dev->bins[c];
    o.modulation = f->modulation == MOD_NFM ? ABG_MOD_NFM : ABG_MOD_AM;
    o.needs_raw_iq = ch->needs_raw_iq;
    o.has_iq_outputs = ch->has_iq_outputs;
    o.dm_dphi = ch->dm_dphi;
    o.alpha = ch->alpha;
    o.ampfactor = f->ampfactor;
    o.squelch_level = f->squelch_level;
    o.squelch_snr_db = f->squelch_snr_db;
    o.lowpass_hz = f->lowpass_hz;
    o.notch_hz = f->notch_hz;
    o.notch_q = f->notch_q;
    o.ctcss_hz = f->ctcss_hz;
    o.afc = ch->afc;
}

extern "C" void* demodulate_b200(void* params) {
    demod_params_t* dp = (demod_params_t*)params;
    const int d0 = dp->device_start, d1 = dp->device_end, nd = d1 - d0;
    device_t* devices = g_b200.devices;
    const int B = g_b200.wave_rate / 8;  // WAVE_BATCH

    // ---- engine set-up: init_demod() + top of demodulate() (:253-266,292-351) ----
    std::vector<abg_device_cfg> dcfg(nd);
    std::vector<std::vector<abg_channel_cfg>> ccfg(nd);
    for (int i = 0; i < nd; i++) {
        device_t* dev = devices + d0 + i;
        ccfg[i].resize(dev->channel_count);
        for (int c = 0; c < dev->channel_count; c++) {
            channel_t* ch = dev->channels + c;
            fill_channel_cfg(dev, c, ch->freqlist + ch->freq_idx, &ccfg[i][c]);
        }
        dcfg[i].sfmt = (int32_t)dev->input->sfmt;
        dcfg[i].fullscale = dev->input->fullscale;
        dcfg[i].sample_rate = dev->input->sample_rate;
        dcfg[i].n_channels = dev->channel_count;
        dcfg[i].channels = ccfg[i].data();
    }
    abg_config cfg;
    cfg.fft_size = (int32_t)g_b200.fft_size;
    cfg.wave_rate = g_b200.wave_rate;
    cfg.fm_demod = g_b200.fm_demod;
    cfg.n_devices = nd;
    cfg.devices = dcfg.data();
    abg_options opt;
    memset(&opt, 0, sizeof(opt));
    opt.cuda_device = -1;
    opt.max_batches_per_run = g_b200.max_batches_per_run > 0 ? g_b200.max_batches_per_run : 2;
    opt.input_capacity_batches = opt.max_batches_per_run + 2;
    abg_engine* eng = nullptr;
    if (abg_create(&cfg, &opt, &eng) != ABG_OK) {
        fatal("Unable to start the B200 demodulation engine");
        return NULL;
    }
    // ---- ingest bridge: pin the input rings in place so that abg_push is asynchronous DMA straight out of the ring the
    // SDR threads fill (ring + wrap tail, input-helpers.cpp:27-36).  Best effort: pageable rings still work. ----
    std::vector<unsigned char*> pinned_rings;
    for (int i = 0; i < nd; i++) {
        input_t* in = devices[d0 + i].input;
        const size_t ring_bytes = in->buf_size + 2 * (size_t)in->bytes_per_sample * g_b200.fft_size;
        if (abg_host_register(in->buffer, ring_bytes) == ABG_OK) pinned_rings.push_back(in->buffer);
    }
    struct Unpin {
        std::vector<unsigned char*>& v;
        ~Unpin() {
            for (unsigned char* p : v) abg_host_unregister(p);
        }
    } unpin{pinned_rings};
    // ---- scan mode: channels with a frequency list (rtl_airband.h:250-252) hand the whole list to the engine; the entry
    // in use follows channel_t.freq_idx, which controller_thread changes (rtl_airband.cpp:117-119) ----
    std::vector<std::vector<int>> scan_idx(nd);  // freq_idx the engine currently uses, -1 = not a scan channel
    for (int i = 0; i < nd; i++) {
        device_t* dev = devices + d0 + i;
        scan_idx[i].assign(dev->channel_count, -1);
        for (int c = 0; c < dev->channel_count; c++) {
            channel_t* ch = dev->channels + c;
            if (ch->freq_count < 2) continue;
            std::vector<abg_channel_cfg> list(ch->freq_count);
            for (int k = 0; k < ch->freq_count; k++) fill_channel_cfg(dev, c, ch->freqlist + k, &list[k]);
            if (abg_scan_configure(eng, i, c, ch->freq_count, list.data()) != ABG_OK) {
                fatal("abg_scan_configure failed");
                abg_destroy(eng);
                return NULL;
            }
            scan_idx[i][c] = 0;
        }
    }
    g_b200.engine_ready = 1;
    std::vector<float> wo, iq;
    std::vector<char> axc;
    bool idle = false;  // the previous pass neither pushed, demodulated nor delivered anything
    while (true) {
        if (g_b200.do_exit) {
            abg_destroy(eng);
            return NULL;
        }
        if (g_b200.devices_running == 0 && idle) {  // :377-381 — but only once everything buffered has been delivered
            g_b200.do_exit = 1;                         // log(LOG_ERR, "All receivers failed, exiting\n") in the reference tree
            continue;
        }
        bool pushed = false;
        std::vector<size_t> new_bufs(nd);
        for (int i = 0; i < nd; i++) {
            device_t* dev = devices + d0 + i;
            input_t* in = dev->input;
            size_t available;
            pthread_mutex_lock(&in->buffer_lock);  // :370-375
            if (in->bufe >= in->bufs)
                available = in->bufe - in->bufs;
            else
                available = in->buf_size - in->bufs + in->bufe;
            pthread_mutex_unlock(&in->buffer_lock);
            if (in->state != INPUT_RUNNING) {  // :383-391
                if (in->state == INPUT_FAILED) {
                    in->state = INPUT_DISABLED;
                    g_b200.devices_running--;
                }
                // whatever is still buffered is demodulated (the reference also drains until `available` runs short)
            }
            const size_t bpc = 2 * (size_t)in->bytes_per_sample;
            const size_t hop_bytes = bpc * (size_t)abg_hop(eng, i);
            // hand over whole hops only (the reference advances bufs hop by hop, :669), at most one batch per visit
            size_t n = std::min(available / hop_bytes, (size_t)B) * hop_bytes;
            if (in->state != INPUT_RUNNING && n == 0 && available >= bpc) n = (available / bpc) * bpc;  // final partial hop at EOF
            size_t local_bufs = in->bufs;
            while (n > 0) {
                const size_t chunk = std::min(n, in->buf_size - local_bufs);  // up to the physical end of the ring
                int rc = abg_push(eng, i, in->buffer + local_bufs, chunk);
                if (rc == ABG_EOVERFLOW) break;  // engine buffer full: demodulate first
                if (rc != ABG_OK) {
                    fatal("abg_push failed");
                    abg_destroy(eng);
                    return NULL;
                }
                // the ring space is released (bufs advanced, not under the lock, like :669) only after the copy has
                // left the ring: with a page-locked ring abg_push is asynchronous
                local_bufs = (local_bufs + chunk) % in->buf_size;
                n -= chunk;
                pushed = true;
            }
            new_bufs[i] = local_bufs;
        }
        for (int i = 0; i < nd; i++) {  // fparms = freqlist + freq_idx, re-read before every batch (:498)
            device_t* dev = devices + d0 + i;
            for (int c = 0; c < dev->channel_count; c++) {
                const int want = dev->channels[c].freq_idx;
                if (scan_idx[i][c] < 0 || want == scan_idx[i][c]) continue;
                if (abg_scan_select(eng, i, c, want) != ABG_OK) {
                    fatal("abg_scan_select failed");
                    abg_destroy(eng);
                    return NULL;
                }
                scan_idx[i][c] = want;
            }
        }
        int produced = abg_run(eng, -1);
        if (produced < 0 && produced != ABG_EOVERFLOW) {
            fatal("abg_run failed");
            abg_destroy(eng);
            return NULL;
        }
        // (after abg_run, so that waiting for the copies overlaps the kernels that were just enqueued)
        if (pushed) {
            if (abg_ingest_sync(eng) != ABG_OK) {
                fatal("abg_ingest_sync failed");
                abg_destroy(eng);
                return NULL;
            }
            for (int i = 0; i < nd; i++) devices[d0 + i].input->bufs = new_bufs[i];
        }
        // ---- deliver finished batches (:621-662 + output.cpp:903-923 hand-shake) ----
        bool delivered = false;
        for (int i = 0; i < nd; i++) {
            device_t* dev = devices + d0 + i;
            while (abg_batches_ready(eng, i) > 0) {
                if (g_b200.wait_for_consumer && dev->waveavail == 1) break;  // offline pacing: do not overrun the output thread
                const int C = dev->channel_count;
                wo.resize((size_t)C * B);
                iq.resize((size_t)C * 2 * B);
                axc.resize(C);
                int rc = abg_fetch_batch(eng, i, wo.data(), iq.data(), axc.data());
                if (rc < 0) {
                    fatal("abg_fetch_batch failed");
                    abg_destroy(eng);
                    return NULL;
                }
                if (rc == 0) break;
                for (int c = 0; c < C; c++) {
                    channel_t* ch = dev->channels + c;
                    memcpy(ch->waveout, wo.data() + (size_t)c * B, sizeof(float) * B);
                    if (ch->has_iq_outputs) memcpy(ch->iq_out, iq.data() + (size_t)c * 2 * B, sizeof(float) * 2 * B);
                    ch->axcindicate = (enum status)axc[c];
                    if (ch->axcindicate != NO_SIGNAL) ch->freqlist[ch->freq_idx].active_counter++;  // :645-647
                }
                if (dev->waveavail == 1)
                    dev->output_overrun_count++;  // :649-652
                else
                    dev->waveavail = 1;
                dp->mp3_signal->send();  // :662
                delivered = true;
                if (g_b200.wait_for_consumer) break;
            }
        }
        idle = !pushed && !delivered && produced <= 0;
        if (idle) {
            bool waiting = false;  // batches held back only because the output thread has not consumed the previous one
            for (int i = 0; i < nd; i++) waiting = waiting || abg_batches_ready(eng, i) > 0;
            if (waiting) idle = false;
            usleep(waiting ? 200 : 10 * 1000);  // SLEEP(10), :398
        }
    }
}

// refresh the Squelch getters of one channel for the stats file / TUI (output.cpp:606-766, rtl_airband.cpp:632-643)
extern "C" ABG_API int b200_refresh_stats(abg_engine* eng, int dev_local, device_t* dev) {
    for (int c = 0; c < dev->channel_count; c++) {
        abg_squelch_stats s;
        int rc = abg_get_stats(eng, dev_local, c, &s);
        if (rc != ABG_OK) return rc;
        freq_t* f = dev->channels[c].freqlist + dev->channels[c].freq_idx;
        f->noise_level = s.noise_level;
        f->signal_level = s.signal_level;
        f->squelch_level_now = s.squelch_level;
        f->open_count = s.open_count;
        f->flappy_count = s.flappy_count;
        f->ctcss_count = s.ctcss_count;
        f->no_ctcss_count = s.no_ctcss_count;
        f->agcavgfast = s.agcavgfast;
        dev->bins[c] = (size_t)s.bin;
    }
    return ABG_OK;
}

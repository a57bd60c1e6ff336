// ISO BMFF (MP4) writer: one AVC or HEVC video track and, optionally, one AAC audio track.
//   ftyp / moov / mdat with stsd(avc1+avcC | hvc1+hvcC | mp4a+esds), stts, stss, stsc, stsz, stco|co64.
//
// Streaming: samples are appended to `mdat` in the file as the encoder delivers them (a 60-minute
// task no longer holds its whole payload in memory); `moov` is written when the task ends.  With
// `-movflags +faststart` it must sit in FRONT of `mdat`: space for it is reserved after `ftyp` from the
// expected number of samples (the rest becomes a `free` box); if the estimate was too small the payload is
// moved up in place, which is what libavformat's second pass does for every file.
//
// Replaces libavformat's `mov` muxer inside the ffmpeg child the reference spawns
// (/root/reference/cmd/consumer.go:376-382; output is always *.mp4, /root/reference/cmd/producer.go:417-425;
// every encode preset carries `-c:a aac -b:a Nk -movflags +faststart`, internal/config/config.go:45-50).
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "host_bits.h"
#include "host_util.h"

namespace vcp {

std::vector<NalRef> split_annexb(const uint8_t* d, size_t n) {
    std::vector<NalRef> out;
    size_t i = 0, start = (size_t)-1;
    while (i + 3 <= n) {
        if (d[i] == 0 && d[i + 1] == 0 && d[i + 2] == 1) {
            if (start != (size_t)-1) {
                size_t e = i;
                while (e > start && d[e - 1] == 0) e--;
                out.push_back({d + start, e - start});
            }
            start = i + 3;
            i += 3;
        } else i++;
    }
    if (start != (size_t)-1 && start <= n) out.push_back({d + start, n - start});
    return out;
}

namespace {

struct Box {
    std::vector<uint8_t> d;
    void u8(uint32_t v) { d.push_back((uint8_t)v); }
    void u16(uint32_t v) { u8(v >> 8); u8(v); }
    void u24(uint32_t v) { u8(v >> 16); u8(v >> 8); u8(v); }
    void u32(uint32_t v) { u8(v >> 24); u8(v >> 16); u8(v >> 8); u8(v); }
    void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
    void tag(const char* t) { d.insert(d.end(), t, t + 4); }
    void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; d.insert(d.end(), q, q + n); }
    void zeros(size_t n) { d.insert(d.end(), n, 0); }
    size_t begin(const char* t) { size_t at = d.size(); u32(0); tag(t); return at; }
    void end(size_t at) {
        uint32_t sz = (uint32_t)(d.size() - at);
        d[at] = sz >> 24; d[at + 1] = sz >> 16; d[at + 2] = sz >> 8; d[at + 3] = sz;
    }
    void full(uint32_t version, uint32_t flags) { u8(version); u24(flags); }
};

void unity_matrix(Box& b) {
    const uint32_t m[9] = {0x00010000, 0, 0, 0, 0x00010000, 0, 0, 0, 0x40000000};
    for (uint32_t v : m) b.u32(v);
}

// HEVCDecoderConfigurationRecord (ISO/IEC 14496-15 8.3.3.1): profile / tier / level copied out of the SPS,
// then one complete array per parameter-set type
void hvcc_box(Box& b, const ParamSets& ps) {
    std::vector<uint8_t> sps;            // SPS without emulation prevention bytes
    for (size_t i = 0, z = 0; i < ps.sps.size(); i++) {
        const uint8_t v = ps.sps[i];
        if (z >= 2 && v == 3) { z = 0; continue; }
        sps.push_back(v);
        z = v == 0 ? z + 1 : 0;
    }
    size_t c = b.begin("hvcC");
    b.u8(1);
    // sps: 2 bytes NAL header, 1 byte vps id / sub layers / nesting, then general profile_tier_level: 12 bytes
    for (int i = 0; i < 12; i++) b.u8(sps.size() > (size_t)(3 + i) ? sps[3 + i] : 0);
    b.u16(0xF000);                       // min_spatial_segmentation_idc 0
    b.u8(0xFC);                          // parallelismType 0 (unknown)
    b.u8(0xFC | 1);                      // chroma_format_idc 4:2:0
    b.u8(0xF8); b.u8(0xF8);              // bit depth luma / chroma minus 8
    b.u16(0);                            // avgFrameRate: unspecified
    b.u8((0 << 6) | (1 << 3) | (1 << 2) | 3);   // constantFrameRate 0, numTemporalLayers 1, temporalIdNested 1, 4-byte lengths
    b.u8(3);
    const std::vector<uint8_t>* arr[3] = {&ps.vps, &ps.sps, &ps.pps};
    const int types[3] = {32, 33, 34};
    for (int i = 0; i < 3; i++) {
        b.u8(0x80 | types[i]);           // array_completeness 1
        b.u16(1);
        b.u16((uint32_t)arr[i]->size()); b.bytes(arr[i]->data(), arr[i]->size());
    }
    b.end(c);
}

// MPEG-4 descriptor: tag, length in the 4-byte 0x80-continued form, payload
void descr(Box& b, int tag, const std::vector<uint8_t>& payload) {
    b.u8(tag);
    const uint32_t n = (uint32_t)payload.size();
    b.u8(0x80 | ((n >> 21) & 0x7f)); b.u8(0x80 | ((n >> 14) & 0x7f)); b.u8(0x80 | ((n >> 7) & 0x7f)); b.u8(n & 0x7f);
    b.bytes(payload.data(), payload.size());
}

// sample-to-chunk runs, chunk offsets
void chunk_tables(Box& b, const std::vector<Mp4Chunk>& chunks, bool co64) {
    size_t a = b.begin("stsc"); b.full(0, 0);
    const size_t cnt_at = b.d.size(); b.u32(0);
    uint32_t runs = 0, prev = 0;
    for (size_t i = 0; i < chunks.size(); i++)
        if (i == 0 || chunks[i].nsamples != prev) { b.u32((uint32_t)i + 1); b.u32(chunks[i].nsamples); b.u32(1); prev = chunks[i].nsamples; runs++; }
    b.d[cnt_at] = runs >> 24; b.d[cnt_at + 1] = runs >> 16; b.d[cnt_at + 2] = runs >> 8; b.d[cnt_at + 3] = runs;
    b.end(a);
    if (co64) { a = b.begin("co64"); b.full(0, 0); b.u32((uint32_t)chunks.size()); for (const auto& c : chunks) b.u64(c.offset); b.end(a); }
    else { a = b.begin("stco"); b.full(0, 0); b.u32((uint32_t)chunks.size()); for (const auto& c : chunks) b.u32((uint32_t)c.offset); b.end(a); }
}

void dinf_box(Box& b) {
    size_t a = b.begin("dinf");
    size_t d2 = b.begin("dref"); b.full(0, 0); b.u32(1);
    size_t u = b.begin("url "); b.full(0, 1); b.end(u);
    b.end(d2); b.end(a);
}

}  // namespace

// moov for the tracks as they stand; every chunk offset is shifted by `shift` (payload moved / moov in front)
std::vector<uint8_t> Mp4Writer::build_moov(uint64_t shift) const {
    const std::vector<uint8_t>&sps = ps_.sps, &pps = ps_.pps;
    const bool hevc = p_.codec == VCPENC_CODEC_HEVC;
    const uint32_t n = (uint32_t)vsize_.size();
    const uint32_t mts = (uint32_t)p_.fps_num, delta = (uint32_t)p_.fps_den;   // media timescale / sample delta
    const uint64_t mdur = (uint64_t)n * delta;
    const uint32_t movie_ts = 1000;
    const uint64_t dur = mts ? mdur * movie_ts / mts : 0;
    const bool have_audio = audio_.present && !asize_.empty();
    uint64_t last = 0;
    for (const auto& c : vchunks_) last = std::max(last, c.offset + shift);
    for (const auto& c : achunks_) last = std::max(last, c.offset + shift);
    const bool co64 = last > 0xFFFFFFFFull;
    auto shifted = [&](const std::vector<Mp4Chunk>& in) { std::vector<Mp4Chunk> o = in; for (auto& c : o) c.offset += shift; return o; };
    Box b;
    size_t moov = b.begin("moov");
    {
        size_t a = b.begin("mvhd"); b.full(0, 0);
        b.u32(0); b.u32(0); b.u32(movie_ts); b.u32((uint32_t)dur);
        b.u32(0x00010000); b.u16(0x0100); b.u16(0); b.u32(0); b.u32(0);
        unity_matrix(b);
        b.zeros(24); b.u32(have_audio ? 3 : 2);
        b.end(a);
    }
    {   // ---- video track ----
        size_t trak = b.begin("trak");
        {
            size_t a = b.begin("tkhd"); b.full(0, 3);
            b.u32(0); b.u32(0); b.u32(1); b.u32(0); b.u32((uint32_t)dur);
            b.u32(0); b.u32(0); b.u16(0); b.u16(0); b.u16(0); b.u16(0);
            unity_matrix(b);
            b.u32((uint32_t)p_.width << 16); b.u32((uint32_t)p_.height << 16);
            b.end(a);
        }
        size_t mdia = b.begin("mdia");
        {
            size_t a = b.begin("mdhd"); b.full(0, 0);
            b.u32(0); b.u32(0); b.u32(mts); b.u32((uint32_t)mdur); b.u16(0x55C4); b.u16(0);
            b.end(a);
            a = b.begin("hdlr"); b.full(0, 0);
            b.u32(0); b.tag("vide"); b.zeros(12); b.bytes("VideoHandler", 13);
            b.end(a);
        }
        size_t minf = b.begin("minf");
        { size_t a = b.begin("vmhd"); b.full(0, 1); b.zeros(8); b.end(a); }
        dinf_box(b);
        size_t stbl = b.begin("stbl");
        {
            size_t a = b.begin("stsd"); b.full(0, 0); b.u32(1);
            size_t e = b.begin(hevc ? "hvc1" : "avc1");
            b.zeros(6); b.u16(1);
            b.zeros(16);
            b.u16((uint32_t)p_.width); b.u16((uint32_t)p_.height);
            b.u32(0x00480000); b.u32(0x00480000); b.u32(0); b.u16(1);
            b.zeros(32);
            b.u16(0x0018); b.u16(0xFFFF);
            if (hevc) hvcc_box(b, ps_);
            else {
                size_t c = b.begin("avcC");
                b.u8(1); b.u8(sps.size() > 1 ? sps[1] : 66); b.u8(sps.size() > 2 ? sps[2] : 0); b.u8(sps.size() > 3 ? sps[3] : 40);
                b.u8(0xFF); b.u8(0xE1); b.u16((uint32_t)sps.size()); b.bytes(sps.data(), sps.size());
                b.u8(1); b.u16((uint32_t)pps.size()); b.bytes(pps.data(), pps.size());
                b.end(c);
            }
            b.end(e); b.end(a);

            a = b.begin("stts"); b.full(0, 0); b.u32(1); b.u32(n); b.u32(delta); b.end(a);
            a = b.begin("stss"); b.full(0, 0);
            b.u32((uint32_t)vsync_.size());
            for (uint32_t i : vsync_) b.u32(i + 1);
            b.end(a);
            a = b.begin("stsz"); b.full(0, 0); b.u32(0); b.u32(n);
            for (uint32_t s : vsize_) b.u32(s);
            b.end(a);
            chunk_tables(b, shifted(vchunks_), co64);
        }
        b.end(stbl); b.end(minf); b.end(mdia); b.end(trak);
    }
    if (have_audio) {   // ---- audio track: AAC in mp4a / esds ----
        const uint32_t ats = (uint32_t)audio_.sample_rate, na = (uint32_t)asize_.size();
        const uint64_t adur_media = (uint64_t)na * (uint32_t)audio_.frame_samples;
        const uint64_t played = adur_media > (uint64_t)audio_.priming ? adur_media - audio_.priming : 0;
        const uint64_t adur = ats ? played * movie_ts / ats : 0;
        size_t trak = b.begin("trak");
        {
            size_t a = b.begin("tkhd"); b.full(0, 3);
            b.u32(0); b.u32(0); b.u32(2); b.u32(0); b.u32((uint32_t)adur);
            b.u32(0); b.u32(0); b.u16(0); b.u16(1); b.u16(0x0100); b.u16(0);
            unity_matrix(b);
            b.u32(0); b.u32(0);
            b.end(a);
        }
        if (audio_.priming > 0) {   // edit list: playback starts after the encoder's priming samples
            size_t e = b.begin("edts");
            size_t a = b.begin("elst"); b.full(0, 0); b.u32(1); b.u32((uint32_t)adur); b.u32((uint32_t)audio_.priming); b.u32(0x00010000); b.end(a);
            b.end(e);
        }
        size_t mdia = b.begin("mdia");
        {
            size_t a = b.begin("mdhd"); b.full(0, 0);
            b.u32(0); b.u32(0); b.u32(ats); b.u32((uint32_t)adur_media); b.u16(0x55C4); b.u16(0);
            b.end(a);
            a = b.begin("hdlr"); b.full(0, 0);
            b.u32(0); b.tag("soun"); b.zeros(12); b.bytes("SoundHandler", 13);
            b.end(a);
        }
        size_t minf = b.begin("minf");
        { size_t a = b.begin("smhd"); b.full(0, 0); b.u16(0); b.u16(0); b.end(a); }
        dinf_box(b);
        size_t stbl = b.begin("stbl");
        {
            size_t a = b.begin("stsd"); b.full(0, 0); b.u32(1);
            size_t e = b.begin("mp4a");
            b.zeros(6); b.u16(1);
            b.zeros(8);
            b.u16((uint32_t)audio_.channels); b.u16(16); b.u16(0); b.u16(0);
            b.u32(ats <= 0xFFFF ? ats << 16 : 0);
            {
                size_t es = b.begin("esds"); b.full(0, 0);
                uint32_t maxsz = 0; uint64_t total = 0;
                for (uint32_t s : asize_) { maxsz = std::max(maxsz, s); total += s; }
                const uint32_t avg = adur_media ? (uint32_t)(total * 8 * ats / adur_media) : (uint32_t)audio_.bitrate;
                Box dsi; descr(dsi, 0x05, audio_.asc);
                Box dcd;
                {
                    std::vector<uint8_t> pl;
                    Box t; t.u8(0x40); t.u8(0x15); t.u24(maxsz ? maxsz : 1536); t.u32(std::max<uint32_t>(avg, (uint32_t)audio_.bitrate)); t.u32(avg);
                    pl = t.d; pl.insert(pl.end(), dsi.d.begin(), dsi.d.end());
                    descr(dcd, 0x04, pl);
                }
                Box sl; descr(sl, 0x06, std::vector<uint8_t>{0x02});
                std::vector<uint8_t> esp = {0x00, 0x02, 0x00};   // ES_ID 2, no flags
                esp.insert(esp.end(), dcd.d.begin(), dcd.d.end());
                esp.insert(esp.end(), sl.d.begin(), sl.d.end());
                descr(b, 0x03, esp);
                b.end(es);
            }
            b.end(e); b.end(a);
            a = b.begin("stts"); b.full(0, 0); b.u32(1); b.u32(na); b.u32((uint32_t)audio_.frame_samples); b.end(a);
            a = b.begin("stsz"); b.full(0, 0); b.u32(0); b.u32(na);
            for (uint32_t s : asize_) b.u32(s);
            b.end(a);
            chunk_tables(b, shifted(achunks_), co64);
        }
        b.end(stbl); b.end(minf); b.end(mdia); b.end(trak);
    }
    b.end(moov);
    return b.d;
}

Mp4Writer::~Mp4Writer() { if (f_) fclose(f_); }

int Mp4Writer::open(const char* path, const vcpenc_params& p, uint64_t expect_vsamples, uint64_t expect_asamples, char* err, size_t errlen) {
    p_ = p; path_ = path;
    f_ = fopen(path, "wb+");
    if (!f_) { set_err(err, errlen, "cannot create %s", path); return VCPENC_E_IO; }
    iobuf_.resize(4 << 20);
    setvbuf(f_, iobuf_.data(), _IOFBF, iobuf_.size());
    Box ftyp;
    size_t a = ftyp.begin("ftyp");
    ftyp.tag("isom"); ftyp.u32(0x200); ftyp.tag("isom"); ftyp.tag("iso2"); ftyp.tag(p.codec == VCPENC_CODEC_HEVC ? "hvc1" : "avc1"); ftyp.tag("mp41");
    ftyp.end(a);
    bool ok = fwrite(ftyp.d.data(), 1, ftyp.d.size(), f_) == ftyp.d.size();
    pos_ = ftyp.d.size();
    moov_at_ = pos_;
    reserve_ = 0;
    if (p.faststart) {
        // moov in front of mdat: 4 B per sample (stsz) + sync samples + one chunk per ~GOP per track + fixed part, with
        // slack; co64 entries assumed.  Unknown length: a modest reserve, the payload is moved up at the end if it was too small.
        const uint64_t nv = expect_vsamples ? expect_vsamples : 16384, na = expect_asamples ? expect_asamples : (expect_vsamples ? 0 : 32768);
        const uint64_t gop = (uint64_t)std::max(1, p.gop);
        reserve_ = 4096 + nv * 4 + (nv / gop + 2) * (4 + 12 + 8) * 2 + na * 4 + (na ? 2048 : 0);
        reserve_ += reserve_ / 8;
        std::vector<uint8_t> z((size_t)reserve_, 0);
        Box fr; fr.u32((uint32_t)reserve_); fr.tag("free");
        memcpy(z.data(), fr.d.data(), 8);
        ok = ok && fwrite(z.data(), 1, z.size(), f_) == z.size();
        pos_ += reserve_;
    }
    mdat_at_ = pos_;
    Box mh; mh.u32(1); mh.tag("mdat"); mh.u64(0);   // 64-bit size, patched when the task ends
    ok = ok && fwrite(mh.d.data(), 1, mh.d.size(), f_) == mh.d.size();
    pos_ += 16;
    if (!ok) { set_err(err, errlen, "short write to %s", path); return VCPENC_E_IO; }
    return VCPENC_OK;
}

bool Mp4Writer::put(int track, const uint8_t* d, size_t n) {
    std::vector<Mp4Chunk>& ch = track ? achunks_ : vchunks_;
    if (cur_track_ != track || ch.empty()) { ch.push_back({pos_, 0}); cur_track_ = track; }
    ch.back().nsamples++;
    pos_ += n;
    return n == 0 || fwrite(d, 1, n, f_) == n;
}

int Mp4Writer::video_access_unit(const uint8_t* annexb, size_t len, bool sync) {
    // parameter sets go to avcC / hvcC; every other NAL unit gets a 4-byte length prefix
    scratch_.clear();
    for (const auto& nal : split_annexb(annexb, len)) {
        if (!nal.n) continue;
        if (ps_.take(p_.codec, nal)) continue;
        const uint32_t k = (uint32_t)nal.n;
        const uint8_t h[4] = {(uint8_t)(k >> 24), (uint8_t)(k >> 16), (uint8_t)(k >> 8), (uint8_t)k};
        scratch_.insert(scratch_.end(), h, h + 4);
        scratch_.insert(scratch_.end(), nal.p, nal.p + nal.n);
    }
    if (sync) vsync_.push_back((uint32_t)vsize_.size());
    vsize_.push_back((uint32_t)scratch_.size());
    if (!put(0, scratch_.data(), scratch_.size())) failed_ = true;
    return failed_ ? VCPENC_E_IO : VCPENC_OK;
}

int Mp4Writer::audio_frame(const uint8_t* d, size_t n) {
    asize_.push_back((uint32_t)n);
    if (!put(1, d, n)) failed_ = true;
    return failed_ ? VCPENC_E_IO : VCPENC_OK;
}

void Mp4Writer::end_chunk() { cur_track_ = -1; }

// move [from, from + len) up by `delta` bytes, back to front
static bool shift_up(FILE* f, uint64_t from, uint64_t len, uint64_t delta) {
    if (fflush(f) != 0) return false;
    const int fd = fileno(f);
    std::vector<uint8_t> buf(8 << 20);
    uint64_t left = len;
    while (left) {
        const size_t n = (size_t)std::min<uint64_t>(left, buf.size());
        const uint64_t at = from + left - n;
        if (pread(fd, buf.data(), n, (off_t)at) != (ssize_t)n) return false;
        if (pwrite(fd, buf.data(), n, (off_t)(at + delta)) != (ssize_t)n) return false;
        left -= n;
    }
    return true;
}

int Mp4Writer::finish(const AudioTrack* audio, char* err, size_t errlen) {
    if (!f_) { set_err(err, errlen, "mp4 writer not open"); return VCPENC_E_INTERNAL; }
    if (audio) { audio_ = *audio; audio_.data.clear(); audio_.sizes.clear(); }
    const bool hevc = p_.codec == VCPENC_CODEC_HEVC;
    if (hevc && ps_.vps.empty()) ps_.vps = make_hevc_vps_nal(p_);
    if (ps_.sps.empty()) ps_.sps = hevc ? make_hevc_sps_nal(p_) : make_sps_nal(p_);
    if (ps_.pps.empty()) ps_.pps = hevc ? make_hevc_pps_nal(p_) : make_pps_nal(p_);
    bool ok = !failed_;
    const uint64_t mdat_len = pos_ - mdat_at_;   // header included
    uint64_t shift = 0;
    std::vector<uint8_t> moov = build_moov(0);
    if (p_.faststart && moov.size() + 8 > reserve_ && moov.size() != reserve_) {
        // the reserve was too small: move the payload up (the moov grows when offsets pass 4 GiB: settle the size first)
        for (int it = 0; it < 3; it++) {
            shift = ((moov.size() + 8 - reserve_) + 4095) & ~(uint64_t)4095;
            std::vector<uint8_t> m2 = build_moov(shift);
            const bool same = m2.size() == moov.size();
            moov.swap(m2);
            if (same) break;
        }
        ok = ok && shift_up(f_, mdat_at_, mdat_len, shift);
        reserve_ += shift; mdat_at_ += shift; pos_ += shift;
    }
    ok = ok && fflush(f_) == 0;
    // mdat size
    Box mh; mh.u32(1); mh.tag("mdat"); mh.u64(mdat_len);
    ok = ok && fseeko(f_, (off_t)mdat_at_, SEEK_SET) == 0 && fwrite(mh.d.data(), 1, 16, f_) == 16;
    if (p_.faststart) {
        ok = ok && fseeko(f_, (off_t)moov_at_, SEEK_SET) == 0 && fwrite(moov.data(), 1, moov.size(), f_) == moov.size();
        const uint64_t rest = reserve_ - moov.size();
        if (rest >= 8) { Box fr; fr.u32((uint32_t)rest); fr.tag("free"); ok = ok && fwrite(fr.d.data(), 1, 8, f_) == 8; }
        else if (rest != 0) ok = false;   // cannot happen: a reserve within 8 bytes of the moov is grown above
    } else {
        ok = ok && fseeko(f_, (off_t)pos_, SEEK_SET) == 0 && fwrite(moov.data(), 1, moov.size(), f_) == moov.size();
    }
    ok = (fclose(f_) == 0) && ok;
    f_ = nullptr;
    if (!ok) { set_err(err, errlen, "short write to %s", path_.c_str()); return VCPENC_E_IO; }
    return VCPENC_OK;
}

void Mp4Writer::abandon() {
    if (f_) { fclose(f_); f_ = nullptr; }
    if (!path_.empty()) remove(path_.c_str());
}

}  // namespace vcp

using namespace vcp;

// Annex-B (with per-frame index) -> MP4, optionally with an AAC track beside it (raw access units back to back in
// `aac`, their sizes in `aac_sizes`; asc = AudioSpecificConfig).  Audio is interleaved GOP by GOP.
extern "C" int vcpenc_mux_mp4_audio(const vcpenc_params* p, const uint8_t* annexb, size_t len, const vcpenc_frame_info* info,
                                    int nframes, const uint8_t* aac, const uint32_t* aac_sizes, int aac_frames, int sample_rate,
                                    int channels, int priming, const uint8_t* asc, int asc_len, const char* path, char* err, size_t errlen) {
    if (!p || !annexb || !info || nframes < 1 || !path) { set_err(err, errlen, "bad arguments"); return VCPENC_E_ARGS; }
    if (aac_frames > 0 && (!aac || !aac_sizes || !asc || asc_len < 2 || sample_rate <= 0 || channels <= 0)) { set_err(err, errlen, "bad audio arguments"); return VCPENC_E_ARGS; }
    for (int i = 0; i < nframes; i++)
        if (info[i].offset + info[i].size > len) { set_err(err, errlen, "frame index out of range"); return VCPENC_E_ARGS; }
    AudioTrack at;
    if (aac_frames > 0) {
        at.present = true; at.sample_rate = sample_rate; at.channels = channels; at.priming = priming; at.frame_samples = 1024;
        at.asc.assign(asc, asc + asc_len);
    }
    Mp4Writer w;
    int rc = w.open(path, *p, (uint64_t)nframes, (uint64_t)std::max(0, aac_frames), err, errlen);
    if (rc) return rc;
    const int gop = std::max(1, p->gop);
    int ai = 0; size_t ao = 0;
    for (int i = 0; i < nframes && !rc; i++) {
        rc = w.video_access_unit(annexb + info[i].offset, info[i].size, info[i].is_idr != 0);
        if ((i + 1) % gop == 0 || i + 1 == nframes) {
            w.end_chunk();
            // audio up to the end of this GOP (all of it after the last one)
            const long long upto = i + 1 == nframes ? aac_frames
                : (p->fps_num > 0 ? (long long)(i + 1) * p->fps_den * sample_rate / ((long long)p->fps_num * 1024) : aac_frames);
            for (; ai < aac_frames && ai < upto && !rc; ai++) { rc = w.audio_frame(aac + ao, aac_sizes[ai]); ao += aac_sizes[ai]; }
            w.end_chunk();
        }
    }
    if (!rc) rc = w.finish(aac_frames > 0 ? &at : nullptr, err, errlen);
    else set_err(err, errlen, "short write to %s", path);
    if (rc) w.abandon();
    return rc;
}

extern "C" int vcpenc_mux_mp4(const vcpenc_params* p, const uint8_t* annexb, size_t len, const vcpenc_frame_info* info,
                              int nframes, const char* path, char* err, size_t errlen) {
    return vcpenc_mux_mp4_audio(p, annexb, len, info, nframes, nullptr, nullptr, 0, 0, 0, 0, nullptr, 0, path, err, errlen);
}

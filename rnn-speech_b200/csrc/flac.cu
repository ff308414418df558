// Host-side FLAC stream decoder (no device work in this file).
//
// The reference decodes audio files inside librosa.load (util/audioprocessor.py:49) through
// audioread / soundfile -> libFLAC for the LibriSpeech corpus its README trains on
// (README.md:60-75, util/dataprocessor.py:207-243 walks *.flac).  None of those libraries is in
// this image, so the container format is decoded here, from the published FLAC format
// specification: STREAMINFO, frame header (with CRC-8), CONSTANT / VERBATIM / FIXED / LPC
// subframes, Rice-coded residuals (both parameter widths, escaped partitions), wasted bits,
// the three stereo decorrelation modes and the frame CRC-16.  Output is interleaved int32
// samples; the int16 -> float32 / mono / 22 050 Hz steps run on the device (resample.cu).
#include "common.cuh"
#include <string.h>
#include <vector>

namespace rs {
namespace {

// MSB-first bit reader over a byte buffer: a 64-bit accumulator holds the next `cnt` bits left-aligned.
struct BitReader {
  const uint8_t* p;
  size_t n, next;         // next: index of the first byte not yet in the accumulator
  uint64_t acc;
  int cnt;
  bool fail;
  BitReader(const uint8_t* data, size_t bytes, size_t start_byte)
      : p(data), n(bytes), next(start_byte < bytes ? start_byte : bytes), acc(0), cnt(0), fail(false) {}
  inline void refill() {
    while (cnt <= 56 && next < n) {
      acc |= (uint64_t)p[next++] << (56 - cnt);
      cnt += 8;
    }
  }
  inline uint64_t take(int k) {             // 1 <= k <= 32
    if (cnt < k) {
      refill();
      if (cnt < k) { fail = true; cnt = 0; acc = 0; return 0; }
    }
    const uint64_t v = acc >> (64 - k);
    acc <<= k;
    cnt -= k;
    return v;
  }
  inline uint32_t bit() { return (uint32_t)take(1); }
  inline uint64_t bits(int k) {             // k <= 64
    if (k <= 0) return 0;
    if (k <= 32) return take(k);
    const uint64_t hi = take(k - 32);
    return (hi << 32) | take(32);
  }
  inline int64_t sbits(int k) {
    if (k == 0) return 0;
    const uint64_t v = bits(k);
    const uint64_t sign = 1ull << (k - 1);
    return (int64_t)((v ^ sign)) - (int64_t)sign;
  }
  inline uint32_t unary() {                 // number of 0 bits before the next 1
    uint32_t q = 0;
    for (;;) {
      if (cnt == 0) {
        refill();
        if (cnt == 0) { fail = true; return q; }
      }
      if (acc == 0) {                       // all `cnt` buffered bits are zero
        q += (uint32_t)cnt;
        cnt = 0;
        continue;
      }
      const int z = __builtin_clzll(acc);   // z < cnt: bits below the valid ones are kept zero
      q += (uint32_t)z;
      acc = z == 63 ? 0 : acc << (z + 1);   // a shift by 64 is undefined
      cnt -= z + 1;
      return q;
    }
  }
  inline void align() {
    const int drop = cnt & 7;
    acc <<= drop;
    cnt -= drop;
  }
  inline size_t byte_pos() const { return next - (size_t)(cnt >> 3); }   // meaningful when byte-aligned
};

uint8_t crc8(const uint8_t* d, size_t n) {          // polynomial x^8 + x^2 + x + 1
  uint8_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int k = 0; k < 8; ++k) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
  }
  return c;
}

struct Crc16Table {                                 // polynomial x^16 + x^15 + x^2 + 1
  uint16_t v[256];
  Crc16Table() {
    for (int i = 0; i < 256; ++i) {
      uint16_t c = (uint16_t)(i << 8);
      for (int k = 0; k < 8; ++k) c = (c & 0x8000) ? (uint16_t)((c << 1) ^ 0x8005) : (uint16_t)(c << 1);
      v[i] = c;
    }
  }
};

uint16_t crc16(const uint8_t* d, size_t n) {
  static const Crc16Table table;                    // initialised once, thread-safe (C++11 static local)
  uint16_t c = 0;
  for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ table.v[((c >> 8) ^ d[i]) & 0xff]);
  return c;
}

bool read_residual(BitReader& br, int64_t* s, int blocksize, int order) {
  const int method = (int)br.bits(2);
  if (method > 1) return false;
  const int pbits = method == 0 ? 4 : 5;
  const int escape = method == 0 ? 15 : 31;
  const int porder = (int)br.bits(4);
  const int nparts = 1 << porder;
  if ((blocksize >> porder) << porder != blocksize && porder > 0) return false;
  int i = order;
  for (int part = 0; part < nparts; ++part) {
    int count = (blocksize >> porder) - (part == 0 ? order : 0);
    if (count < 0) return false;
    const int param = (int)br.bits(pbits);
    if (param == escape) {
      const int raw = (int)br.bits(5);
      for (int k = 0; k < count; ++k) s[i++] = br.sbits(raw);
    } else {
      for (int k = 0; k < count; ++k) {
        const uint64_t q = br.unary();
        const uint64_t u = (q << param) | (param ? br.bits(param) : 0);
        s[i++] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
      }
    }
    if (br.fail) return false;
  }
  return i == blocksize;
}

bool read_subframe(BitReader& br, int64_t* s, int blocksize, int bps) {
  if (br.bit()) return false;                              // padding bit
  const int type = (int)br.bits(6);
  int wasted = 0;
  if (br.bit()) wasted = (int)br.unary() + 1;
  bps -= wasted;
  if (bps < 1) return false;
  if (type == 0) {                                         // CONSTANT
    const int64_t v = br.sbits(bps);
    for (int i = 0; i < blocksize; ++i) s[i] = v;
  } else if (type == 1) {                                  // VERBATIM
    for (int i = 0; i < blocksize; ++i) s[i] = br.sbits(bps);
  } else if (type >= 8 && type <= 12) {                    // FIXED, order type - 8
    const int order = type - 8;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) s[i] = br.sbits(bps);
    if (!read_residual(br, s, blocksize, order)) return false;
    switch (order) {
      case 1: for (int i = 1; i < blocksize; ++i) s[i] += s[i - 1]; break;
      case 2: for (int i = 2; i < blocksize; ++i) s[i] += 2 * s[i - 1] - s[i - 2]; break;
      case 3: for (int i = 3; i < blocksize; ++i) s[i] += 3 * s[i - 1] - 3 * s[i - 2] + s[i - 3]; break;
      case 4: for (int i = 4; i < blocksize; ++i) s[i] += 4 * s[i - 1] - 6 * s[i - 2] + 4 * s[i - 3] - s[i - 4]; break;
      default: break;
    }
  } else if (type >= 32) {                                 // LPC, order type - 31
    const int order = type - 31;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) s[i] = br.sbits(bps);
    const int precision = (int)br.bits(4) + 1;
    if (precision == 16) return false;
    const int shift = (int)br.sbits(5);
    if (shift < 0) return false;
    int64_t coef[32];
    for (int j = 0; j < order; ++j) coef[j] = br.sbits(precision);
    if (!read_residual(br, s, blocksize, order)) return false;
    for (int i = order; i < blocksize; ++i) {
      int64_t acc = 0;
      for (int j = 0; j < order; ++j) acc += coef[j] * s[i - 1 - j];
      s[i] += acc >> shift;
    }
  } else {
    return false;                                          // reserved subframe type
  }
  if (wasted)
    for (int i = 0; i < blocksize; ++i) s[i] = s[i] * ((int64_t)1 << wasted);
  return !br.fail;
}

}  // namespace
}  // namespace rs

using namespace rs;

// data / nbytes: a whole .flac file in host memory.  out: interleaved int32 samples, capacity out_capacity VALUES
// (frames * channels), or NULL to parse only (out_capacity >= 0: trust STREAMINFO's count; < 0: walk and count).  Returns RS_OK and fills sample_rate, channels,
// bits_per_sample, total_frames (decoded count when out != NULL or the header does not state it) and md5[16]
// (STREAMINFO's signature of the unencoded audio, all zero when the encoder did not write one).
extern "C" int rs_flac_decode_host(const uint8_t* data, size_t nbytes, int32_t* out, int64_t out_capacity,
                                   int* sample_rate, int* channels, int* bits_per_sample, int64_t* total_frames,
                                   uint8_t* md5) {
  RS_REQUIRE(data != nullptr && nbytes >= 42, RS_ERR_INVALID, "rs_flac_decode_host: buffer too small for a FLAC stream");
  size_t pos = 0;
  if (memcmp(data, "ID3", 3) == 0 && nbytes > 10) {        // an ID3v2 tag in front of the stream marker
    const size_t sz = ((size_t)(data[6] & 0x7f) << 21) | ((size_t)(data[7] & 0x7f) << 14) | ((size_t)(data[8] & 0x7f) << 7) |
                      (size_t)(data[9] & 0x7f);
    pos = 10 + sz;
  }
  RS_REQUIRE(pos + 4 <= nbytes && memcmp(data + pos, "fLaC", 4) == 0, RS_ERR_INVALID,
             "rs_flac_decode_host: no fLaC stream marker");
  pos += 4;
  int sr = 0, nch = 0, bps = 0, max_block = 0;
  int64_t total = 0;
  bool have_info = false, last = false;
  while (!last) {
    RS_REQUIRE(pos + 4 <= nbytes, RS_ERR_INVALID, "rs_flac_decode_host: truncated metadata");
    last = (data[pos] & 0x80) != 0;
    const int type = data[pos] & 0x7f;
    const size_t len = ((size_t)data[pos + 1] << 16) | ((size_t)data[pos + 2] << 8) | data[pos + 3];
    pos += 4;
    RS_REQUIRE(pos + len <= nbytes, RS_ERR_INVALID, "rs_flac_decode_host: truncated metadata block");
    if (type == 0) {
      RS_REQUIRE(len >= 34, RS_ERR_INVALID, "rs_flac_decode_host: short STREAMINFO");
      const uint8_t* q = data + pos;
      max_block = (q[2] << 8) | q[3];
      sr = (q[10] << 12) | (q[11] << 4) | (q[12] >> 4);
      nch = ((q[12] >> 1) & 7) + 1;
      bps = (((q[12] & 1) << 4) | (q[13] >> 4)) + 1;
      total = ((int64_t)(q[13] & 0x0f) << 32) | ((int64_t)q[14] << 24) | ((int64_t)q[15] << 16) | ((int64_t)q[16] << 8) | q[17];
      if (md5) memcpy(md5, q + 18, 16);
      have_info = true;
    }
    pos += len;
  }
  RS_REQUIRE(have_info && sr > 0 && nch >= 1 && nch <= 8 && bps >= 4 && bps <= 32 && max_block >= 16, RS_ERR_INVALID,
             "rs_flac_decode_host: bad STREAMINFO (sr %d, channels %d, bits %d, block %d)", sr, nch, bps, max_block);
  if (sample_rate) *sample_rate = sr;
  if (channels) *channels = nch;
  if (bits_per_sample) *bits_per_sample = bps;
  if (out == nullptr && total > 0 && out_capacity >= 0) {      // out_capacity < 0: walk the frames and count anyway
    if (total_frames) *total_frames = total;
    return RS_OK;
  }

  // per-thread scratch, kept across calls (a fresh 0.5-4 MB vector per file means an mmap / munmap pair per call, which
  // serialises decoder threads on the process's address-space lock)
  static thread_local std::vector<int64_t> buf;
  if (buf.size() < (size_t)nch * 65536) buf.resize((size_t)nch * 65536);
  int64_t done = 0;                                        // frames decoded so far
  while (pos + 2 <= nbytes) {
    if (!(data[pos] == 0xFF && (data[pos + 1] & 0xFE) == 0xF8)) {
      // trailing bytes that are not a frame (padding, tags): stop
      bool rest_zero = true;
      for (size_t i = pos; i < nbytes && rest_zero; ++i) rest_zero = data[i] == 0;
      RS_REQUIRE(rest_zero || done > 0, RS_ERR_INVALID, "rs_flac_decode_host: lost frame sync at byte %zu", pos);
      break;
    }
    const size_t frame_start = pos;
    BitReader br(data, nbytes, pos);
    br.bits(15);
    br.bit();                                              // blocking strategy: only changes what the coded number means
    const int bs_code = (int)br.bits(4), sr_code = (int)br.bits(4);
    const int ch_code = (int)br.bits(4), ss_code = (int)br.bits(3);
    RS_REQUIRE(br.bit() == 0, RS_ERR_INVALID, "rs_flac_decode_host: reserved bit set in frame header at byte %zu", pos);
    {                                                      // UTF-8 style coded frame / sample number
      const uint32_t first = (uint32_t)br.bits(8);
      int extra = 0;
      if (first & 0x80) {
        uint32_t m = 0x40;
        while (first & m) { ++extra; m >>= 1; }
        RS_REQUIRE(extra >= 1 && extra <= 6, RS_ERR_INVALID, "rs_flac_decode_host: bad coded number at byte %zu", pos);
      }
      for (int i = 0; i < extra; ++i) br.bits(8);
    }
    int blocksize;
    if (bs_code == 1) blocksize = 192;
    else if (bs_code >= 2 && bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = (int)br.bits(8) + 1;
    else if (bs_code == 7) blocksize = (int)br.bits(16) + 1;
    else if (bs_code >= 8) blocksize = 256 << (bs_code - 8);
    else { rs::set_error("rs_flac_decode_host: reserved block size code at byte %zu", pos); return RS_ERR_INVALID; }
    if (sr_code == 12) br.bits(8);
    else if (sr_code == 13 || sr_code == 14) br.bits(16);
    RS_REQUIRE(sr_code != 15, RS_ERR_INVALID, "rs_flac_decode_host: invalid sample rate code at byte %zu", pos);
    static const int ss_table[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    const int fbps = ss_code == 0 ? bps : ss_table[ss_code];
    RS_REQUIRE(fbps > 0, RS_ERR_INVALID, "rs_flac_decode_host: reserved sample size code at byte %zu", pos);
    const size_t hdr_end = br.byte_pos();
    RS_REQUIRE(!br.fail && hdr_end < nbytes, RS_ERR_INVALID, "rs_flac_decode_host: truncated frame header");
    const uint8_t want8 = (uint8_t)br.bits(8);
    RS_REQUIRE(crc8(data + frame_start, hdr_end - frame_start) == want8, RS_ERR_INVALID,
               "rs_flac_decode_host: frame header CRC-8 mismatch at byte %zu", frame_start);
    int fch;
    if (ch_code < 8) fch = ch_code + 1;
    else if (ch_code <= 10) fch = 2;
    else { rs::set_error("rs_flac_decode_host: reserved channel assignment at byte %zu", pos); return RS_ERR_INVALID; }
    RS_REQUIRE(fch == nch && blocksize <= 65536, RS_ERR_INVALID, "rs_flac_decode_host: frame has %d channels, stream %d",
               fch, nch);
    for (int c = 0; c < nch; ++c) {
      const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
      if (!read_subframe(br, buf.data() + (size_t)c * 65536, blocksize, fbps + (side ? 1 : 0))) {
        rs::set_error("rs_flac_decode_host: bad subframe (channel %d) in the frame at byte %zu", c, frame_start);
        return RS_ERR_INVALID;
      }
    }
    br.align();
    const size_t body_end = br.byte_pos();
    RS_REQUIRE(body_end + 2 <= nbytes, RS_ERR_INVALID, "rs_flac_decode_host: truncated frame at byte %zu", frame_start);
    const uint16_t want16 = (uint16_t)((data[body_end] << 8) | data[body_end + 1]);
    RS_REQUIRE(crc16(data + frame_start, body_end - frame_start) == want16, RS_ERR_INVALID,
               "rs_flac_decode_host: frame CRC-16 mismatch at byte %zu", frame_start);
    pos = body_end + 2;
    int64_t* c0 = buf.data();
    int64_t* c1 = buf.data() + 65536;
    if (ch_code == 8) for (int i = 0; i < blocksize; ++i) c1[i] = c0[i] - c1[i];                  // left, side
    else if (ch_code == 9) for (int i = 0; i < blocksize; ++i) c0[i] = c0[i] + c1[i];             // side, right
    else if (ch_code == 10)                                                                        // mid, side
      for (int i = 0; i < blocksize; ++i) {
        const int64_t side = c1[i];
        const int64_t mid = (c0[i] * 2) | (side & 1);
        c0[i] = (mid + side) >> 1;
        c1[i] = (mid - side) >> 1;
      }
    if (out) {
      RS_REQUIRE((done + blocksize) * nch <= out_capacity, RS_ERR_WORKSPACE,
                 "rs_flac_decode_host: output holds %lld values, stream has more", (long long)out_capacity);
      for (int c = 0; c < nch; ++c) {
        const int64_t* src = buf.data() + (size_t)c * 65536;
        int32_t* dst = out + done * nch + c;
        for (int i = 0; i < blocksize; ++i) dst[(size_t)i * nch] = (int32_t)src[i];
      }
    }
    done += blocksize;
  }
  RS_REQUIRE(total == 0 || done == total, RS_ERR_INVALID, "rs_flac_decode_host: decoded %lld frames, STREAMINFO says %lld",
             (long long)done, (long long)total);
  if (total_frames) *total_frames = done;
  return RS_OK;
}

// Host half of model loading: dnn.bin → one relocatable blob that is uploaded (or NCCL-broadcast)
// as is.  No CUDA in this file; it works without a GPU (fdnn_pack).
//
// Follows the behaviour of the reference loader and quantizer (paths under /root/reference):
//   file layout / byte order      src/cpp/float_dnn.cc:18-69, 166-212 (4-byte big-endian words)
//   layer-0 input padding to ×4   src/cpp/float_dnn.cc:32-33, 61-66, 76-83
//   int8 weight quantization      src/cpp/dnn.cc:460-509 (+ absMax :148-160)
//   sigmoid lookup table          src/cpp/dnn.cc:100-115, src/cpp/dnn.h:35-42
// including its quirks: the multiplier is an integer-valued float, only the lower clip is live,
// and the float→char conversion wraps modulo 256 the way x86 cvttss2si + truncation does.

#include "fdnn_internal.h"

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cctype>
#include <cstring>
#include <memory>

#include "../../include/fdnn.h"

namespace fdnn {

namespace {

thread_local std::string g_error;

struct FileBytes {
  std::vector<uint8_t> data;
  size_t pos = 0;
  bool short_read = false;

  uint32_t word() {
    if (pos + 4 > data.size()) {
      short_read = true;
      return 0;
    }
    const uint8_t *b = data.data() + pos;
    pos += 4;
    return (uint32_t(b[0]) << 24) | (uint32_t(b[1]) << 16) | (uint32_t(b[2]) << 8) | uint32_t(b[3]);
  }
  // bulk big-endian fp32 → native
  bool floats(float *dst, size_t count) {
    if (count > (data.size() - pos) / 4) {
      short_read = true;
      return false;
    }
    const uint8_t *b = data.data() + pos;
    for (size_t i = 0; i < count; ++i, b += 4) {
      uint32_t u = (uint32_t(b[0]) << 24) | (uint32_t(b[1]) << 16) | (uint32_t(b[2]) << 8) | uint32_t(b[3]);
      std::memcpy(dst + i, &u, 4);
    }
    pos += count * 4;
    return true;
  }
};

int read_file(const char *path, std::vector<uint8_t> &out) {
  FILE *f = std::fopen(path, "rb");
  if (!f) {
    set_error(std::string("cannot open ") + path);
    return FDNN_EIO;
  }
  std::fseek(f, 0, SEEK_END);
  long size = std::ftell(f);
  std::rewind(f);
  if (size < 0) {
    std::fclose(f);
    set_error(std::string("cannot size ") + path);
    return FDNN_EIO;
  }
  out.resize(size_t(size));
  size_t got = size ? std::fread(out.data(), 1, size_t(size), f) : 0;
  std::fclose(f);
  if (got != size_t(size)) {
    set_error(std::string("short read on ") + path);
    return FDNN_EIO;
  }
  return FDNN_OK;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// What static_cast<int>(float) does on x86-64 (cvttss2si): values outside int32 and NaN give
// INT_MIN.  The reference reaches this through static_cast<char>(round(f·mult)) (dnn.cc:499).
inline int x86_float_to_int(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
  return int(v);
}

struct FloatLayer {
  int in = 0, out = 0, in_padded = 0;
  std::vector<float> w;  // [out][in_padded]
  std::vector<float> bias;
};

// Largest |w| after clipping to ±cutoff over the whole layer (dnn.cc:148-160, 469-476).
float clipped_abs_max(const FloatLayer &l, float cutoff) {
  float best = -3.402823466e+38f;
  for (float v : l.w) {
    if (v < -cutoff) v = -cutoff;
    if (v > cutoff) v = cutoff;
    float a = std::fabs(v);
    if (a > best) best = a;
  }
  return best;
}

// The 3-instruction division q = s·r; e = fma(−q, c, s); q' = fma(e, r, q) with r = RN(1/c) is
// checked against IEEE s / c for every numerator the layer can produce (|s| ≤ K/2 · 32768, as a
// float) — cheap because numerators are integers.  Needs hardware fma on the host to be exact
// and fast; without it the layer simply keeps the IEEE division on the device.
__attribute__((target("fma"))) bool verify_fast_div_fma(float coeff, float rcp, int K) {
  const int64_t limit = int64_t(K / 2) * 32768;
  auto same = [&](float s) {
    float q = s * rcp;
    float e = __builtin_fmaf(-q, coeff, s);
    float q2 = __builtin_fmaf(e, rcp, q);
    float ref = s / coeff;
    return std::memcmp(&q2, &ref, 4) == 0;
  };
  // every float value that (float)int32 can take in [0, limit]; the negative side is symmetric
  int64_t s = 0;
  while (s <= limit) {
    if (!same(float(s))) return false;
    int64_t step = 1;
    if (s >= (int64_t(1) << 24)) step = 2;
    if (s >= (int64_t(1) << 25)) step = 4;
    if (s >= (int64_t(1) << 26)) step = 8;
    if (s >= (int64_t(1) << 27)) step = 16;
    s += step;
  }
  return true;
}

bool verify_fast_div(float coeff, float rcp, int K) {
  if (!__builtin_cpu_supports("fma")) return false;
  struct Memo {
    float coeff;
    int K;
    bool ok;
  };
  static thread_local std::vector<Memo> memo;
  for (const Memo &m : memo)
    if (m.coeff == coeff && m.K == K) return m.ok;
  bool ok = verify_fast_div_fma(coeff, rcp, K);
  memo.push_back({coeff, K, ok});
  return ok;
}

}  // namespace

void set_error(const std::string &msg) { g_error = msg; }
const char *get_error() { return g_error.c_str(); }

void build_reference_lut(uint8_t out[1280]) {
  for (int i = -kLutHalf; i < kLutHalf; ++i) {
    float k = float(i) / 100.0f;
    float s = 1.0f / (1 + expf(-k));
    out[i + kLutHalf] = uint8_t(x86_float_to_int(roundf(s * 255.0f)));
  }
}

int pack_model(const char *path, float cutoff, std::vector<uint8_t> &blob) {
  if (!path) {
    set_error("null path");
    return FDNN_EINVAL;
  }
  if (!(cutoff > 0)) {  // QuantizedDnn.java:55-57 rejects these before going native
    set_error("weight cutoff must be positive");
    return FDNN_EINVAL;
  }
  FileBytes fb;
  if (int rc = read_file(path, fb.data)) return rc;

  const int layer_count = int(fb.word());
  if (fb.short_read || layer_count < 3 || layer_count > 4096) {
    set_error("not a usable dnn.bin: need at least 3 layers (one fp32 input layer and two int8 layers)");
    return FDNN_EFORMAT;
  }
  std::vector<FloatLayer> layers(size_t(layer_count), FloatLayer{});
  for (int j = 0; j < layer_count; ++j) {
    FloatLayer &l = layers[size_t(j)];
    l.in = int(fb.word());
    l.out = int(fb.word());
    if (fb.short_read || l.in <= 0 || l.out <= 0 || size_t(l.in) > fb.data.size() || size_t(l.out) > fb.data.size()) {
      set_error("dnn.bin: bad layer header at layer " + std::to_string(j));
      return fb.short_read ? FDNN_EIO : FDNN_EFORMAT;
    }
    l.in_padded = j == 0 ? round_up(l.in, 4) : l.in;
    if (size_t(l.in) * size_t(l.out) > (fb.data.size() - fb.pos) / 4) {
      set_error("dnn.bin truncated inside layer " + std::to_string(j));
      return FDNN_EIO;
    }
    l.w.assign(size_t(l.out) * size_t(l.in_padded), 0.0f);
    for (int o = 0; o < l.out; ++o) fb.floats(l.w.data() + size_t(o) * size_t(l.in_padded), size_t(l.in));
    l.bias.resize(size_t(l.out));
    if (!fb.floats(l.bias.data(), size_t(l.out))) {
      set_error("dnn.bin truncated inside layer " + std::to_string(j));
      return FDNN_EIO;
    }
  }
  const int in_file = layers[0].in, in_dim = layers[0].in_padded, H = layers[0].out;
  std::vector<float> shift(size_t(in_dim), 0.0f), scale(size_t(in_dim), 0.0f);
  if (!fb.floats(shift.data(), size_t(in_file)) || !fb.floats(scale.data(), size_t(in_file))) {
    set_error("dnn.bin truncated in shift/scale");
    return FDNN_EIO;
  }

  // Constraints the reference relies on without checking (dnn.cc:199, 254, 331; README.md:10,69).
  if (H % 16 != 0) {
    set_error("hidden width must be a multiple of 16");
    return FDNN_EFORMAT;
  }
  for (int j = 1; j < layer_count; ++j) {
    if (layers[size_t(j)].in != layers[size_t(j - 1)].out) {
      set_error("layer " + std::to_string(j) + " input width does not match previous layer");
      return FDNN_EFORMAT;
    }
    if (j < layer_count - 1 && layers[size_t(j)].out != H) {
      set_error("all hidden layers must have the same width");
      return FDNN_EFORMAT;
    }
  }
  if (H / 2 > 65536) {
    set_error("hidden width too large");
    return FDNN_EFORMAT;
  }

  const int nq = layer_count - 1;
  std::vector<BlobQLayer> qmeta(size_t(nq), BlobQLayer{});
  std::vector<std::vector<int8_t>> qw(size_t(nq), std::vector<int8_t>{});
  using FixVariants = std::array<std::vector<FixEntry>, kFixVariants>;
  using PtrVariants = std::array<std::vector<uint32_t>, kFixVariants>;
  std::vector<FixVariants> fix(size_t(nq), FixVariants{});
  std::vector<PtrVariants> fix_ptr(size_t(nq), PtrVariants{});

  for (int q = 0; q < nq; ++q) {
    const FloatLayer &l = layers[size_t(q + 1)];
    BlobQLayer &m = qmeta[size_t(q)];
    m.nodes = l.out;
    m.inputs = l.in;
    const float max = clipped_abs_max(l, cutoff);
    m.multiplier = roundf(127.0f / max);
    m.coeff = m.multiplier * 255.0f;
    m.rcp_coeff = 1.0f / m.coeff;
    std::vector<int8_t> &w8 = qw[size_t(q)];
    w8.resize(l.w.size());
    const float lo = -cutoff;
    for (size_t i = 0; i < l.w.size(); ++i) {
      float f = l.w[i];
      if (f < lo) f = lo;  // only the lower clip is live in the reference (dnn.cc:493-498)
      w8[i] = int8_t(uint8_t(x86_float_to_int(roundf(f * m.multiplier)) & 0xff));
    }
    m.fast_div = (std::isfinite(m.coeff) && m.coeff != 0.0f && verify_fast_div(m.coeff, m.rcp_coeff, l.in)) ? 1u : 0u;

    // saturation risk lists, one per tile width G, ordered by (node / G, K block, node, pair)
    const int K = l.in, pairs = K / 2;
    m.k_blocks = uint32_t((K + kFixKBlock - 1) / kFixKBlock);
    auto risky = [](int a, int b) {
      int pos = (a > 0 ? a : 0) + (b > 0 ? b : 0), neg = (a < 0 ? a : 0) + (b < 0 ? b : 0);
      return pos >= 129 || neg <= -129;
    };
    std::vector<FixEntry> all;  // node-major
    for (int n = 0; n < l.out; ++n) {
      const int8_t *row = w8.data() + size_t(n) * size_t(K);
      for (int p = 0; p < pairs; ++p) {
        const int a = row[2 * p], b = row[2 * p + 1];
        if (risky(a, b)) all.push_back(FixEntry{uint32_t(p) | (uint32_t(uint8_t(a)) << 16) | (uint32_t(uint8_t(b)) << 24), uint32_t(n)});
      }
    }
    m.n_fix = uint32_t(all.size());
    for (int v = 0; v < kFixVariants; ++v) {
      const int G = kFixGroups[v];
      const size_t buckets = size_t((l.out + G - 1) / G) * size_t(m.k_blocks);
      auto bucket_of = [&](const FixEntry &e) { return size_t(e.node / uint32_t(G)) * m.k_blocks + size_t(2 * (e.pair_w & 0xffffu) / kFixKBlock); };
      std::vector<uint32_t> &ptr = fix_ptr[size_t(q)][size_t(v)];
      ptr.assign(buckets + 1, 0u);
      for (const FixEntry &e : all) ++ptr[bucket_of(e) + 1];
      for (size_t i = 0; i < buckets; ++i) ptr[i + 1] += ptr[i];
      std::vector<uint32_t> cursor(ptr.begin(), ptr.end() - 1);
      std::vector<FixEntry> &ent = fix[size_t(q)][size_t(v)];
      ent.resize(all.size());
      for (const FixEntry &e : all) ent[cursor[bucket_of(e)]++] = e;
      // Inside a bucket the order is free.  The scan warps read the activation pairs of four
      // consecutive entries (positions 4i … 4i+3 of the bucket) with one shared-memory instruction,
      // which is conflict-free when the four pairs sit at different word offsets of their 16-byte
      // chunk, i.e. differ in (pair >> 1) & 3 — so deal the entries out round-robin by that class.
      std::vector<FixEntry> cls[4], mixed;
      for (size_t b = 0; b < buckets; ++b) {
        const uint32_t lo = ptr[b], hi = ptr[b + 1];
        if (hi - lo < 2) continue;
        for (auto &c : cls) c.clear();
        for (uint32_t e = lo; e < hi; ++e) cls[((ent[e].pair_w & 0xffffu) >> 1) & 3u].push_back(ent[e]);
        mixed.clear();
        size_t taken[4] = {0, 0, 0, 0};
        while (mixed.size() < size_t(hi - lo)) {
          int order[4] = {0, 1, 2, 3};  // fullest class first, so the short ones run out last
          std::sort(order, order + 4, [&](int a, int c) { return cls[a].size() - taken[a] > cls[c].size() - taken[c]; });
          for (int k : order)
            if (taken[k] < cls[k].size()) mixed.push_back(cls[k][taken[k]++]);
        }
        std::copy(mixed.begin(), mixed.end(), ent.begin() + lo);
      }
    }
  }

  // ---- lay the blob out ----------------------------------------------------------------------
  size_t off = align_up(sizeof(BlobHeader), kBlobAlign);
  BlobHeader hdr{};
  hdr.magic = kBlobMagic;
  hdr.version = kBlobVersion;
  hdr.in_dim = in_dim;
  hdr.in_dim_file = in_file;
  hdr.hidden = H;
  hdr.out_dim = layers.back().out;
  hdr.n_qlayers = nq;
  hdr.cutoff = cutoff;
  auto reserve = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, kBlobAlign);
    return uint64_t(at);
  };
  hdr.off_qlayers = reserve(sizeof(BlobQLayer) * size_t(nq));
  hdr.off_lut = reserve(kLut2Padded);
  hdr.off_shift = reserve(sizeof(float) * size_t(in_dim));
  hdr.off_scale = reserve(sizeof(float) * size_t(in_dim));
  hdr.off_bias0 = reserve(sizeof(float) * size_t(H));
  hdr.off_w0 = reserve(sizeof(float) * layers[0].w.size());
  for (int q = 0; q < nq; ++q) {
    BlobQLayer &m = qmeta[size_t(q)];
    m.off_bias = reserve(sizeof(float) * size_t(m.nodes));
    for (int v = 0; v < kFixVariants; ++v) {
      m.off_fix_ptr[v] = reserve(sizeof(uint32_t) * fix_ptr[size_t(q)][size_t(v)].size());
      m.off_fix_ent[v] = reserve(sizeof(FixEntry) * std::max<size_t>(fix[size_t(q)][size_t(v)].size(), 1));
    }
    m.off_w = reserve(qw[size_t(q)].size());
  }
  hdr.total_size = off;

  blob.assign(off, 0);
  uint8_t *base = blob.data();
  std::memcpy(base, &hdr, sizeof(hdr));
  std::memcpy(base + hdr.off_qlayers, qmeta.data(), sizeof(BlobQLayer) * size_t(nq));
  {
    uint8_t lut[1280];
    build_reference_lut(lut);
    uint8_t *lut2 = base + hdr.off_lut;
    for (int v = -kLut2Center; v <= kLut2Center; ++v) {
      const int mag = ((v < 0 ? -v : v) + 1) >> 1;
      const int k = v < 0 ? -mag : mag;
      lut2[v + kLut2Center] = k <= -kLutHalf ? uint8_t(0) : (k >= kLutHalf ? uint8_t(255) : lut[k + kLutHalf]);  // dnn.h:35-42
    }
  }
  std::memcpy(base + hdr.off_shift, shift.data(), sizeof(float) * shift.size());
  std::memcpy(base + hdr.off_scale, scale.data(), sizeof(float) * scale.size());
  std::memcpy(base + hdr.off_bias0, layers[0].bias.data(), sizeof(float) * size_t(H));
  std::memcpy(base + hdr.off_w0, layers[0].w.data(), sizeof(float) * layers[0].w.size());
  for (int q = 0; q < nq; ++q) {
    const BlobQLayer &m = qmeta[size_t(q)];
    std::memcpy(base + m.off_bias, layers[size_t(q + 1)].bias.data(), sizeof(float) * size_t(m.nodes));
    for (int v = 0; v < kFixVariants; ++v) {
      const auto &ptr = fix_ptr[size_t(q)][size_t(v)];
      const auto &ent = fix[size_t(q)][size_t(v)];
      std::memcpy(base + m.off_fix_ptr[v], ptr.data(), sizeof(uint32_t) * ptr.size());
      if (!ent.empty()) std::memcpy(base + m.off_fix_ent[v], ent.data(), sizeof(FixEntry) * ent.size());
    }
    std::memcpy(base + m.off_w, qw[size_t(q)].data(), qw[size_t(q)].size());
  }
  return FDNN_OK;
}

int validate_blob(const uint8_t *blob, size_t size) {
  if (!blob || size < sizeof(BlobHeader)) {
    set_error("blob too small");
    return FDNN_EFORMAT;
  }
  BlobHeader h;
  std::memcpy(&h, blob, sizeof(h));
  if (h.magic != kBlobMagic || h.version != kBlobVersion) {
    set_error("not a fast-dnn model blob (magic/version mismatch)");
    return FDNN_EFORMAT;
  }
  if (h.total_size != size) {
    set_error("blob size does not match its header");
    return FDNN_EFORMAT;
  }
  auto inside = [&](uint64_t off, uint64_t bytes) { return off % kBlobAlign == 0 && off <= size && bytes <= size - off; };
  if (h.in_dim <= 0 || h.in_dim % 4 || h.hidden <= 0 || h.hidden % 16 || h.out_dim <= 0 || h.n_qlayers < 2) {
    set_error("blob header holds an unusable topology");
    return FDNN_EFORMAT;
  }
  if (!inside(h.off_qlayers, sizeof(BlobQLayer) * uint64_t(h.n_qlayers)) || !inside(h.off_lut, kLut2Padded) ||
      !inside(h.off_shift, 4ull * uint64_t(h.in_dim)) || !inside(h.off_scale, 4ull * uint64_t(h.in_dim)) ||
      !inside(h.off_bias0, 4ull * uint64_t(h.hidden)) || !inside(h.off_w0, 4ull * uint64_t(h.hidden) * uint64_t(h.in_dim))) {
    set_error("blob section out of bounds");
    return FDNN_EFORMAT;
  }
  std::vector<BlobQLayer> q(size_t(h.n_qlayers), BlobQLayer{});
  std::memcpy(q.data(), blob + h.off_qlayers, sizeof(BlobQLayer) * q.size());
  int expect_in = h.hidden;
  for (int i = 0; i < h.n_qlayers; ++i) {
    const BlobQLayer &m = q[size_t(i)];
    bool last = i == h.n_qlayers - 1;
    if (m.inputs != expect_in || m.nodes <= 0 || (!last && m.nodes != h.hidden) || (last && m.nodes != h.out_dim) ||
        m.k_blocks != uint32_t((m.inputs + kFixKBlock - 1) / kFixKBlock)) {
      set_error("blob int8 layer " + std::to_string(i) + " has inconsistent dimensions");
      return FDNN_EFORMAT;
    }
    if (!inside(m.off_w, uint64_t(m.nodes) * uint64_t(m.inputs)) || !inside(m.off_bias, 4ull * uint64_t(m.nodes))) {
      set_error("blob int8 layer " + std::to_string(i) + " section out of bounds");
      return FDNN_EFORMAT;
    }
    for (int v = 0; v < kFixVariants; ++v) {
      const uint32_t G = uint32_t(kFixGroups[v]);
      const uint64_t buckets = uint64_t((uint32_t(m.nodes) + G - 1) / G) * uint64_t(m.k_blocks);
      if (!inside(m.off_fix_ptr[v], 4ull * (buckets + 1)) || !inside(m.off_fix_ent[v], sizeof(FixEntry) * uint64_t(m.n_fix))) {
        set_error("blob int8 layer " + std::to_string(i) + " risk list out of bounds");
        return FDNN_EFORMAT;
      }
      const uint32_t *ptr = reinterpret_cast<const uint32_t *>(blob + m.off_fix_ptr[v]);
      const FixEntry *ent = reinterpret_cast<const FixEntry *>(blob + m.off_fix_ent[v]);
      bool ok = ptr[0] == 0 && ptr[buckets] == m.n_fix;
      for (uint64_t bkt = 0; ok && bkt < buckets; ++bkt) {
        ok = ptr[bkt] <= ptr[bkt + 1] && ptr[bkt + 1] <= m.n_fix;
        const uint32_t g = uint32_t(bkt / m.k_blocks), kb = uint32_t(bkt % m.k_blocks);
        for (uint32_t e = ptr[bkt]; ok && e < ptr[bkt + 1]; ++e) {
          const uint32_t p = ent[e].pair_w & 0xffffu;
          ok = ent[e].node < uint32_t(m.nodes) && ent[e].node / G == g && int(p) < m.inputs / 2 && 2 * p / kFixKBlock == kb;
        }
      }
      if (!ok) {
        set_error("blob int8 layer " + std::to_string(i) + " has a corrupt risk list");
        return FDNN_EFORMAT;
      }
    }
    expect_in = m.nodes;
  }
  return FDNN_OK;
}

// ---- offline tooling either side of the hot path -----------------------

namespace {

void put_be32(std::vector<uint8_t> &out, uint32_t v) {
  out.push_back(uint8_t(v >> 24));
  out.push_back(uint8_t(v >> 16));
  out.push_back(uint8_t(v >> 8));
  out.push_back(uint8_t(v));
}
void put_be_float(std::vector<uint8_t> &out, float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  put_be32(out, u);
}
int write_file(const char *path, const std::vector<uint8_t> &bytes) {
  FILE *f = std::fopen(path, "wb");
  if (!f) {
    set_error(std::string("cannot create ") + path);
    return FDNN_EIO;
  }
  const size_t put = bytes.empty() ? 0 : std::fwrite(bytes.data(), 1, bytes.size(), f);
  const bool ok = std::fclose(f) == 0 && put == bytes.size();
  if (!ok) {
    set_error(std::string("short write on ") + path);
    return FDNN_EIO;
  }
  return FDNN_OK;
}

}  // namespace

// FeedForwardNetwork.align(inputAlignment, hiddenAlignment) + saveBinary
// (src/java/suskun/nn/FeedForwardNetwork.java:50-58, 226-235, 264-281, 331-340): pad the input
// width (and shift/scale) to a multiple of `input_alignment`, every hidden width to a multiple of
// `hidden_alignment`, with zero weights and zero bias; the output layer keeps its width.  The
// reference's README lists doing this on the C++ side as a TODO (README.md:76).
int align_dnn_bin(const char *in_path, const char *out_path, int input_alignment, int hidden_alignment) {
  if (!in_path || !out_path || input_alignment <= 0 || hidden_alignment <= 0) {
    set_error("bad argument to fdnn_align_dnn_bin");
    return FDNN_EINVAL;
  }
  // alignments are SIMD / tile widths (the reference uses 4 and 16, FeedForwardNetwork.java:50-58); a huge one would pad a
  // small network into terabytes of zeros, one push_back at a time
  constexpr int kMaxAlignment = 4096;
  if (input_alignment > kMaxAlignment || hidden_alignment > kMaxAlignment) {
    set_error("alignment above " + std::to_string(kMaxAlignment));
    return FDNN_EINVAL;
  }
  FileBytes fb;
  if (int rc = read_file(in_path, fb.data)) return rc;
  const int layer_count = int(fb.word());
  if (fb.short_read || layer_count < 1 || layer_count > 4096) {
    set_error("not a dnn.bin file");
    return FDNN_EFORMAT;
  }
  std::vector<FloatLayer> layers(size_t(layer_count), FloatLayer{});
  for (int j = 0; j < layer_count; ++j) {
    FloatLayer &l = layers[size_t(j)];
    l.in = int(fb.word());
    l.out = int(fb.word());
    if (fb.short_read || l.in <= 0 || l.out <= 0 || size_t(l.in) * size_t(l.out) > (fb.data.size() - fb.pos) / 4) {
      set_error("dnn.bin truncated or corrupt at layer " + std::to_string(j));
      return fb.short_read ? FDNN_EIO : FDNN_EFORMAT;
    }
    if (j > 0 && l.in != layers[size_t(j - 1)].out) {
      set_error("layer " + std::to_string(j) + " input width does not match previous layer");
      return FDNN_EFORMAT;
    }
    l.w.resize(size_t(l.in) * size_t(l.out));
    l.bias.resize(size_t(l.out));
    if (!fb.floats(l.w.data(), l.w.size()) || !fb.floats(l.bias.data(), l.bias.size())) {
      set_error("dnn.bin truncated inside layer " + std::to_string(j));
      return FDNN_EIO;
    }
  }
  const int in0 = layers[0].in;
  std::vector<float> shift(size_t(in0), 0.0f), scale(size_t(in0), 0.0f);
  if (!fb.floats(shift.data(), shift.size()) || !fb.floats(scale.data(), scale.size())) {
    set_error("dnn.bin truncated in shift/scale");
    return FDNN_EIO;
  }
  // size of the aligned file, known before a byte of it is built: an allocation that cannot succeed fails here, at once
  // (std::bad_alloc → FDNN_ENOMEM at the C ABI)
  size_t total = 4;
  for (int j = 0; j < layer_count; ++j) {
    const FloatLayer &l = layers[size_t(j)];
    const size_t in_pad = size_t(round_up(l.in, j == 0 ? input_alignment : hidden_alignment));
    const size_t out_pad = size_t(j == layer_count - 1 ? l.out : round_up(l.out, hidden_alignment));
    total += 8 + 4 * (in_pad * out_pad + out_pad);
  }
  total += 8 * size_t(round_up(in0, input_alignment));
  std::vector<uint8_t> out;
  out.reserve(total);
  put_be32(out, uint32_t(layer_count));
  for (int j = 0; j < layer_count; ++j) {
    const FloatLayer &l = layers[size_t(j)];
    const int in_pad = round_up(l.in, j == 0 ? input_alignment : hidden_alignment);
    const int out_pad = j == layer_count - 1 ? l.out : round_up(l.out, hidden_alignment);
    put_be32(out, uint32_t(in_pad));
    put_be32(out, uint32_t(out_pad));
    for (int o = 0; o < out_pad; ++o)
      for (int i = 0; i < in_pad; ++i) put_be_float(out, (o < l.out && i < l.in) ? l.w[size_t(o) * size_t(l.in) + size_t(i)] : 0.0f);
    for (int o = 0; o < out_pad; ++o) put_be_float(out, o < l.out ? l.bias[size_t(o)] : 0.0f);
  }
  const int in_pad0 = round_up(in0, input_alignment);
  for (int i = 0; i < in_pad0; ++i) put_be_float(out, i < in0 ? shift[size_t(i)] : 0.0f);
  for (int i = 0; i < in_pad0; ++i) put_be_float(out, i < in0 ? scale[size_t(i)] : 0.0f);
  return write_file(out_path, out);
}

// Kaldi nnet1 text model + feature-transform text → dnn.bin (unaligned; run align_dnn_bin afterwards).
// Follows FeedForwardNetwork.loadFromTextFile / loadLayersFromTextFile (FeedForwardNetwork.java:86-119,
// 159-207): an "<AffineTransform> out in" line announces a layer; lines that start with '<' or consist of a
// single bracket are skipped; the next `out` lines are the weight rows (brackets stripped), one more line is
// the bias.  The transform file is searched for "[ … ]" blocks: three blocks = <Splice> (dropped), shift,
// scale; two blocks = shift, scale; both must be as wide as the first layer's input.
namespace {

std::string trim(const std::string &s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
  while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
  return s.substr(a, b - a);
}

// floats of a line with '[' and ']' removed; false on a token that is not a number
bool parse_floats(const std::string &line, std::vector<float> &out) {
  out.clear();
  std::string clean;
  clean.reserve(line.size());
  for (char c : line) clean.push_back((c == '[' || c == ']') ? ' ' : c);
  const char *p = clean.c_str();
  for (;;) {
    while (*p && std::isspace(static_cast<unsigned char>(*p))) ++p;
    if (!*p) return true;
    char *end = nullptr;
    const float v = std::strtof(p, &end);
    if (end == p) return false;
    out.push_back(v);
    p = end;
  }
}

bool next_line(const std::vector<uint8_t> &text, size_t &pos, std::string &line) {
  if (pos >= text.size()) return false;
  size_t e = pos;
  while (e < text.size() && text[e] != '\n') ++e;
  line.assign(reinterpret_cast<const char *>(text.data()) + pos, e - pos);
  if (!line.empty() && line.back() == '\r') line.pop_back();
  pos = e + 1;
  return true;
}

}  // namespace

int import_kaldi_nnet1(const char *nnet_path, const char *transform_path, const char *out_path) {
  if (!nnet_path || !transform_path || !out_path) {
    set_error("bad argument to fdnn_import_kaldi_nnet1");
    return FDNN_EINVAL;
  }
  std::vector<uint8_t> text;
  if (int rc = read_file(nnet_path, text)) return rc;
  std::vector<FloatLayer> layers;
  size_t pos = 0;
  std::string line;
  int nodes = -1, inputs = -1;
  std::vector<float> row;
  while (next_line(text, pos, line)) {
    line = trim(line);
    if (line.empty()) continue;
    if (line.compare(0, 17, "<AffineTransform>") == 0) {
      std::vector<float> dims;
      if (!parse_floats(line.substr(17), dims) || dims.size() < 2 || dims[0] < 1 || dims[1] < 1 || dims[0] > 1e7f || dims[1] > 1e7f) {
        set_error("malformed <AffineTransform> line in " + std::string(nnet_path));
        return FDNN_EFORMAT;
      }
      nodes = int(dims[0]);
      inputs = int(dims[1]);
    }
    if (nodes == -1 || line[0] == '<' || line == "[" || line == "]") continue;
    FloatLayer l;
    l.in = inputs;
    l.out = nodes;
    l.w.resize(size_t(nodes) * size_t(inputs));
    l.bias.resize(size_t(nodes));
    for (int i = 0; i <= nodes; ++i) {
      if (i > 0 && !next_line(text, pos, line)) {
        set_error("nnet1 text ends inside layer " + std::to_string(layers.size()));
        return FDNN_EFORMAT;
      }
      const size_t want = size_t(i < nodes ? inputs : nodes);
      if (!parse_floats(line, row) || row.size() < want) {
        set_error("layer " + std::to_string(layers.size()) + ", row " + std::to_string(i) + ": expected " + std::to_string(want) + " numbers");
        return FDNN_EFORMAT;
      }
      std::memcpy(i < nodes ? l.w.data() + size_t(i) * size_t(inputs) : l.bias.data(), row.data(), want * sizeof(float));
    }
    if (!layers.empty() && layers.back().out != l.in) {
      set_error("layer " + std::to_string(layers.size()) + " input width does not match previous layer");
      return FDNN_EFORMAT;
    }
    layers.push_back(std::move(l));
  }
  if (layers.empty()) {
    set_error("no <AffineTransform> layers in " + std::string(nnet_path));
    return FDNN_EFORMAT;
  }
  // feature transform: every "[ … ]" block of the whole file
  if (int rc = read_file(transform_path, text)) return rc;
  std::vector<std::vector<float>> blocks;
  for (size_t i = 0; i < text.size(); ++i) {
    if (text[i] != '[') continue;
    size_t e = i + 1;
    while (e < text.size() && text[e] != ']') ++e;
    if (e == text.size()) break;
    std::vector<float> v;
    std::string body(reinterpret_cast<const char *>(text.data()) + i + 1, e - i - 1);
    for (char &c : body)
      if (c == '\n' || c == '\r') c = ' ';
    if (!parse_floats(body, v)) {
      set_error("non-numeric token in a [ ] block of " + std::string(transform_path));
      return FDNN_EFORMAT;
    }
    blocks.push_back(std::move(v));
    i = e;
  }
  if (blocks.size() == 3) blocks.erase(blocks.begin());  // <Splice>
  if (blocks.size() != 2) {
    set_error("unexpected feature transformation vector count: " + std::to_string(blocks.size()));
    return FDNN_EFORMAT;
  }
  const size_t in0 = size_t(layers[0].in);
  if (blocks[0].size() != in0 || blocks[1].size() != in0) {
    set_error("shift/scale vectors (" + std::to_string(blocks[0].size()) + ", " + std::to_string(blocks[1].size()) + ") do not match the input dimension " +
              std::to_string(in0));
    return FDNN_EFORMAT;
  }
  std::vector<uint8_t> out;
  put_be32(out, uint32_t(layers.size()));
  for (const FloatLayer &l : layers) {
    put_be32(out, uint32_t(l.in));
    put_be32(out, uint32_t(l.out));
    for (float f : l.w) put_be_float(out, f);
    for (float f : l.bias) put_be_float(out, f);
  }
  for (float f : blocks[0]) put_be_float(out, f);
  for (float f : blocks[1]) put_be_float(out, f);
  return write_file(out_path, out);
}

// Feature matrix: big-endian int32 frames, int32 dimension, fp32 rows (BatchData.java:80-91, 107-139;
// float_dnn.cc:85-105).  Only the `frames` rows the header announces are read (the shipped
// data/16khz.bin carries one extra row).
int read_feature_bin(const char *path, int *frames, int *dim, std::vector<float> &data) {
  FileBytes fb;
  if (int rc = read_file(path, fb.data)) return rc;
  const int n = int(fb.word()), d = int(fb.word());
  if (fb.short_read || n < 0 || d <= 0 || size_t(n) * size_t(d) > (fb.data.size() - fb.pos) / 4) {
    set_error(std::string("not a usable feature file: ") + path);
    return fb.short_read ? FDNN_EIO : FDNN_EFORMAT;
  }
  data.resize(size_t(n) * size_t(d));
  fb.floats(data.data(), data.size());
  *frames = n;
  *dim = d;
  return FDNN_OK;
}

int write_feature_bin(const char *path, const float *data, int frames, int dim) {
  if (!path || (!data && frames > 0) || frames < 0 || dim <= 0) {
    set_error("bad argument to fdnn_feature_bin_write");
    return FDNN_EINVAL;
  }
  std::vector<uint8_t> out;
  out.reserve(8 + size_t(frames) * size_t(dim) * 4);
  put_be32(out, uint32_t(frames));
  put_be32(out, uint32_t(dim));
  for (size_t i = 0; i < size_t(frames) * size_t(dim); ++i) put_be_float(out, data[i]);
  return write_file(path, out);
}

// BatchData::dumpToFile(…, binary = true) (float_dnn.cc:128-164): NATIVE-endian uint32 n, uint32 d, fp32 rows.
int write_output_dump(const char *path, const float *data, int frames, int dim) {
  if (!path || (!data && frames > 0) || frames < 0 || dim <= 0) {
    set_error("bad argument to fdnn_output_dump_write");
    return FDNN_EINVAL;
  }
  std::vector<uint8_t> out(8 + size_t(frames) * size_t(dim) * 4);
  const uint32_t hdr[2] = {uint32_t(frames), uint32_t(dim)};
  std::memcpy(out.data(), hdr, 8);
  if (frames > 0) std::memcpy(out.data() + 8, data, size_t(frames) * size_t(dim) * 4);
  return write_file(path, out);
}

}  // namespace fdnn

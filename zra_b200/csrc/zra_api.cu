// zra_api.cu — the zra:: C++ API and the Zra* C shim of zra-b200 (host side).
//
// Mirrors the reference's operator surface for the hot path (source/zra.cpp:18-625): same names,
// argument meaning, status codes and documented quirks, with the codec work done by the CUDA
// kernels in this directory. Every entry point takes HOST buffers, stages them through the
// calling thread's GpuContext and fails loudly (ZStdError) when no CUDA device is usable —
// there is no CPU codec in this library.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <optional>

#include "../../include/zra.h"
#include "../../include/zra.hpp"
#include "gpu_context.h"
#include "host_ops.h"
#include "zra_format.h"

namespace zra {
  using namespace zrab;

  // The reference's ZCCtx / ZDCtx wrap a zstd context; ours name the calling thread's GPU context.
  class ZCCtx {
   public:
    i8 level{};
    bool checksum{true};
  };
  // (ZDCtx also carries FullDecompressor's read-ahead: the class layout of zra.hpp is the drop-in contract and has no
  // room for it)
  class ZDCtx {
   public:
    bool aheadValid{false};
    size_t aheadCur{0}, aheadLast{0};   // seek-table range [cur, last] the read-ahead covers
    u64 aheadA{0}, aheadB{0};           // its compressed byte range
    int aheadBuf{0};                    // which pinned staging buffer holds it
    uint64_t aheadGen{0};               // that buffer's hand-out count after the read-ahead (another reader on this thread bumps it)
    int nextBuf{0};                     // buffer of the next call that reads for itself
  };

  struct Entry {
    u8 bytes[5];
  };

  namespace {
    u64 entry_get(const u8* table, size_t index) { return get_le(table + kEntrySize * index, 5); }
    void entry_put(u8* table, size_t index, u64 v) { put_le(table + kEntrySize * index, v, 5); }

    const char* zstd_error_string(int code) {
      switch (code) {
        case 0: return "No error detected";
        case 1: return "Error (generic)";
        case 10: return "Unknown frame descriptor";
        case 12: return "Version not supported";
        case 14: return "Unsupported frame parameter";
        case 16: return "Frame requires too much memory for decoding";
        case 20: return "Corrupted block detected";
        case 22: return "Restored data doesn't match checksum";
        case 30: return "Dictionary is corrupted";
        case 32: return "Dictionary mismatch";
        case 34: return "Cannot create Dictionary from provided samples";
        case 40: return "Unsupported parameter";
        case 42: return "Parameter is out of bound";
        case 44: return "tableLog requires too much memory : unsupported";
        case 46: return "Unsupported max Symbol Value : too large";
        case 48: return "Specified maxSymbolValue is too small";
        case 60: return "Operation not authorized at current processing stage";
        case 62: return "Context should be init first";
        case 64: return "Allocation error : not enough memory";
        case 66: return "workSpace buffer is not large enough";
        case 70: return "Destination buffer is too small";
        case 72: return "Src size is incorrect";
        case 74: return "Operation on NULL destination buffer";
        case 100: return "Frame index is too large";
        case 102: return "An I/O error occurred when reading/seeking";
        case 104: return "Destination buffer is wrong";
        default: return "Unspecified error code";
      }
    }

    [[noreturn]] void throw_gpu(GpuContext* g) {
      fprintf(stderr, "%s\n", g->last_error().c_str());
      throw Exception(StatusCode::ZStdError, 1 /* GENERIC */);
    }

    GpuContext* gpu() {
      GpuContext* g = default_context();
      if (!g->ok()) throw_gpu(g);
      g->bind();
      return g;
    }

    void raise(const OpStatus& s, GpuContext* g) {
      if (s.cuda) throw_gpu(g);
      if (s.zra) throw Exception(static_cast<StatusCode>(s.zra), s.zstd);
    }

    ArchiveInfo info_of(const Header& h) {
      ArchiveInfo a;
      a.headerSize = h.size;
      a.tableSize = h.seekTableSize / (u32)kEntrySize;
      a.frameSize = h.frameSize;
      a.metaSize = h.metaSize;
      a.uncompressedSize = h.uncompressedSize;
      a.frames = a.tableSize ? a.tableSize - 1 : 0;
      return a;
    }
  }  // namespace

  // ------------------------------------------------------------------ errors (source/zra.cpp:46-82)
  Exception::Exception(StatusCode code, int zstdCode) : code(code), zstdCode(zstdCode) {}

  const char* Exception::GetExceptionString(StatusCode code) {
    switch (code) {
      case StatusCode::Success: return "The operation was successful";
      case StatusCode::ZStdError: return "An error was returned by ZStandard";
      case StatusCode::ZraVersionLow: return "This archive was compressed using a newer version of ZRA";
      case StatusCode::HeaderInvalid: return "The header in the supplied buffer was invalid";
      case StatusCode::HeaderIncomplete: return "The header hasn't been fully written before being accessed";
      case StatusCode::OutOfBoundsAccess: return "The specified offset and size are past the data contained within the buffer";
      case StatusCode::OutputBufferTooSmall: return "The output buffer is too small to contain the output";
      case StatusCode::CompressedSizeTooLarge: return "The compressed output's size exceeds the maximum limit";
      case StatusCode::InputFrameSizeMismatch: return "The input size is not divisible by the frame size and nor is it the final frame";
    }
    return "An unknown error has occurred";
  }

  const char* Exception::what() const noexcept {
    if (code != StatusCode::ZStdError) return GetExceptionString(code);
    // the reference builds this into a function-static string; thread_local keeps the lifetime
    // guarantee ("valid until the next call") without the data race
    thread_local std::string reason;
    reason = std::string(GetExceptionString(code)) + ": " + zstd_error_string(zstdCode);
    return reason.c_str();
  }

  u16 GetVersion() { return kZraVersion; }

  // ------------------------------------------------------------------ Header (source/zra.cpp:141-187)
  Header::Header(const std::function<void(size_t, size_t, void*)>& readFunction) : readFunction(readFunction) {
    u8 raw[kFixedHeaderSize];
    readFunction(0, sizeof(raw), raw);
    FixedHeaderFields f = parse_fixed_header(raw);
    if (f.magic != kZraMagic || f.version > GetVersion()) throw Exception(StatusCode::HeaderInvalid);
    version = f.version;
    size = f.headerSize + 8;
    uncompressedSize = f.uncompressedSize;
    frameSize = f.frameSize;
    metaOffset = (u32)kFixedHeaderSize;
    metaSize = f.metaSize;
    seekTableOffset = metaOffset + metaSize;
    seekTableSize = f.tableSize * (u32)kEntrySize;
    if (f.version != 1) throw Exception(StatusCode::ZraVersionLow);
  }

  // The view is captured BY VALUE (the reference captures the caller's BufferView by reference and
  // dangles for temporaries); the `>=` bound is the reference's and is kept.
  Header::Header(const BufferView& buffer)
      : Header([view = buffer](size_t offset, size_t readSize, void* out) {
          if (offset + readSize >= view.size) throw Exception(StatusCode::OutOfBoundsAccess);
          std::memcpy(out, view.data + offset, readSize);
        }) {
    if (buffer.size < size) throw Exception(StatusCode::OutOfBoundsAccess);
  }

  Buffer Header::GetSeekTable() const {
    Buffer table(seekTableSize);
    readFunction(seekTableOffset, seekTableSize, table.data());
    return table;
  }

  void Header::GetMetadata(const BufferView& buffer) const { readFunction(metaOffset, metaSize, buffer.data); }

  Buffer Header::GetMetadata() const {
    Buffer meta(metaSize);
    GetMetadata(meta);
    return meta;
  }

  // ------------------------------------------------------------------ sizes (source/zra.cpp:189-192)
  size_t GetOutputBufferSize(size_t inputSize, u32 frameSize, u32 metaSize) {
    u32 table = table_entries(inputSize, frameSize);
    return kFixedHeaderSize + metaSize + kEntrySize * (size_t)table + zstd_compress_bound(frameSize) * (size_t)(table - 1);
  }

  // ------------------------------------------------------------------ compression (source/zra.cpp:194-241)
  size_t CompressBuffer(const BufferView& input, const BufferView& output, i8 compressionLevel, u32 frameSize, bool checksum,
                        const BufferView& meta) {
    u32 table = table_entries(input.size, frameSize);
    size_t need = kFixedHeaderSize + kEntrySize * (size_t)table + zstd_compress_bound(frameSize) * (size_t)(table - 1);
    if (output.size < need) throw Exception(StatusCode::OutputBufferTooSmall);
    GpuContext* g = gpu();
    size_t written = 0;
    raise(host_compress_buffer(g, input.data, input.size, output.data, output.size, &written, compressionLevel, frameSize, checksum,
                               meta.data, meta.size),
          g);
    return written;
  }

  Buffer CompressBuffer(const BufferView& buffer, i8 compressionLevel, u32 frameSize, bool checksum, const BufferView& meta) {
    Buffer output(GetOutputBufferSize(buffer.size, frameSize));
    output.resize(CompressBuffer(buffer, output, compressionLevel, frameSize, checksum, meta));
    output.shrink_to_fit();
    return output;
  }

  // ------------------------------------------------------------------ decompression (source/zra.cpp:243-302)
  namespace {
    // The frame-parallel decoder reads the seek table (the reference's serial decoder never does): the table must lie
    // where the fixed header says, inside the header. Checked before any GPU work.
    void check_table_geometry(const Header& h) {
      const ArchiveInfo info = info_of(h);
      if (kFixedHeaderSize + (u64)info.metaSize + kEntrySize * (u64)info.tableSize != info.headerSize) throw Exception(StatusCode::HeaderInvalid);
    }
  }  // namespace

  void DecompressBuffer(const BufferView& input, const BufferView& output) {
    Header header(input);
    if (output.size < header.uncompressedSize) throw Exception(StatusCode::OutputBufferTooSmall);
    check_table_geometry(header);
    GpuContext* g = gpu();
    raise(host_decompress_archive(g, input.data, input.size, info_of(header), output.data), g);
  }

  Buffer DecompressBuffer(const BufferView& buffer) {
    // like the reference, the size is taken from the raw header bytes before validation
    Buffer output(buffer.size >= 26 ? get_le(buffer.data + 18, 8) : 0);
    DecompressBuffer(buffer, output);
    return output;
  }

  void DecompressRA(const BufferView& input, const BufferView& output, size_t offset, size_t size) {
    Header header(input);
    // `>=`: the reference's in-memory entry point cannot reach the last byte
    if (offset + size >= header.uncompressedSize) throw Exception(StatusCode::OutOfBoundsAccess);
    if (output.size < size) throw Exception(StatusCode::OutputBufferTooSmall);
    check_table_geometry(header);
    GpuContext* g = gpu();
    raise(host_decompress_range(g, input.data, input.size, info_of(header), offset, size, output.data), g);
  }

  Buffer DecompressRA(const BufferView& buffer, size_t offset, size_t size) {
    Buffer output(size);
    DecompressRA(buffer, output, offset, size);
    return output;
  }

  // ------------------------------------------------------------------ Compressor (source/zra.cpp:304-365)
  Compressor::Compressor(size_t size, i8 compressionLevel, u32 frameSize, bool checksum, const BufferView& meta)
      : ctx(std::make_shared<ZCCtx>()),
        frameSize(frameSize),
        tableSize(table_entries(size, frameSize)),
        header(kFixedHeaderSize + meta.size + kEntrySize * (size_t)tableSize),
        entry(reinterpret_cast<Entry*>(header.data() + kFixedHeaderSize + meta.size)) {
    ctx->level = compressionLevel;
    ctx->checksum = checksum;
    write_fixed_header(header.data(), size, tableSize, frameSize, (u32)meta.size);
    if (meta.data) std::memcpy(header.data() + kFixedHeaderSize, meta.data, meta.size);
  }

  size_t Compressor::GetOutputBufferSize(size_t inputSize) const {
    return zstd_compress_bound(frameSize) * (inputSize / frameSize + ((inputSize % frameSize) ? 1 : 0));
  }

  size_t Compressor::Compress(const BufferView& input, const BufferView& output) {
    if (output.size < GetOutputBufferSize(input.size)) throw Exception(StatusCode::OutputBufferTooSmall);
    // the reference measures the entry index from the start of the metadata section (its
    // pointer arithmetic ignores meta.size); same arithmetic here
    auto entryIndex = static_cast<size_t>(reinterpret_cast<u8*>(entry) - (header.data() + kFixedHeaderSize)) / kEntrySize;
    if (input.size % frameSize && (entryIndex + (input.size / frameSize) + 2) < tableSize)
      throw Exception(StatusCode::InputFrameSizeMismatch);

    size_t frames = input.size / frameSize + ((input.size % frameSize) ? 1 : 0);
    size_t produced = 0;
    if (frames) {
      GpuContext* g = gpu();
      std::vector<u64> sizes(frames);
      raise(host_compress_frames(g, input.data, input.size, frameSize, ctx->level, ctx->checksum, output.data, output.size,
                                 sizes.data(), &produced),
            g);
      u8* e = reinterpret_cast<u8*>(entry);
      for (size_t i = 0; i < frames; i++) {
        put_le(e, outputOffset, 5);
        e += kEntrySize;
        outputOffset += sizes[i];
      }
      entry = reinterpret_cast<Entry*>(e);
      // a short final frame permanently shrinks frameSize, as in the reference (zra.cpp:330)
      if (input.size % frameSize) frameSize = (u32)(input.size % frameSize);
    }
    if (outputOffset >= kMaxCompressedSize) throw Exception(StatusCode::CompressedSizeTooLarge);

    if (reinterpret_cast<u8*>(entry) == header.data() + header.size() - kEntrySize) {
      put_le(reinterpret_cast<u8*>(entry), outputOffset, 5);
      entry = reinterpret_cast<Entry*>(reinterpret_cast<u8*>(entry) + kEntrySize);
      put_le(header.data() + 14, header_hash_host(header.data(), header.size()), 4);
    }
    return produced;
  }

  void Compressor::Compress(const BufferView& input, Buffer& output) {
    output.resize(GetOutputBufferSize(input.size));
    output.resize(Compress(input, BufferView(output)));
  }

  const Buffer& Compressor::GetHeader() {
    if (reinterpret_cast<u8*>(entry) == header.data() + header.size()) return header;
    throw Exception(StatusCode::HeaderIncomplete);
  }

  size_t Compressor::GetHeaderSize() { return header.size(); }

  // ------------------------------------------------------------------ Decompressor (source/zra.cpp:367-424)
  Decompressor::Decompressor(const std::function<void(size_t, size_t, void*)>& readFunction, size_t maxCacheSize)
      : ctx(std::make_shared<ZDCtx>()),
        readFunction(readFunction),
        header(readFunction),
        seekTable(header.GetSeekTable()),
        maxCacheSize(maxCacheSize) {}

  void Decompressor::Decompress(size_t offset, size_t size, const BufferView& output) {
    if (offset + size > header.uncompressedSize) throw Exception(StatusCode::OutOfBoundsAccess);
    if (output.size < size) throw Exception(StatusCode::OutputBufferTooSmall);
    if (!header.frameSize) throw Exception(StatusCode::HeaderInvalid);

    // same frame-range arithmetic and the same single read-callback call as the reference
    u64 q = offset / header.frameSize, r = offset % header.frameSize;
    u64 q2 = (r + size) / header.frameSize, r2 = (r + size) % header.frameSize;
    u64 first = q, last = q + q2 + (r2 ? 1 : 0);
    if (last >= seekTable.size() / kEntrySize) throw Exception(StatusCode::OutOfBoundsAccess);  // table shorter than the geometry
    u64 a = entry_get(seekTable.data(), first), b = entry_get(seekTable.data(), last);
    // the seek table is untrusted input: every entry of the range must lie inside it, in order, before any size is
    // derived from it (a wrapped size would be a huge allocation, a wild upload range or an out-of-bounds kernel read)
    if (b < a) throw Exception(StatusCode::ZStdError, 72);  // srcSize_wrong
    for (u64 f = first; f < last; f++) {
      const u64 fa = entry_get(seekTable.data(), f), fb = entry_get(seekTable.data(), f + 1);
      if (fa < a || fb < fa || fb > b || fb - fa > 0xFFFFFFFFull) throw Exception(StatusCode::ZStdError, 72);
    }
    size_t compressedSize = b - a;

    // The reference reads into `cache` (or a one-off buffer above maxCacheSize, zra.cpp:383-389); here the callback —
    // still exactly one call, on the caller's thread — fills page-locked staging owned by the GPU context, so the
    // upload runs at link speed and no host buffer is zero-filled per request. `cache` stays for the class layout.
    GpuContext* g = gpu();
    u8* input = g->pinned_stage(compressedSize);
    std::optional<Buffer> fallback;
    if (!input) { fallback.emplace(compressedSize); input = fallback->data(); }
    readFunction(header.size + a, compressedSize, input);
    if (first == last) return;

    std::vector<HostFrame> frames(last - first);
    for (u64 f = first; f < last; f++) {
      HostFrame& d = frames[f - first];
      u64 fa = entry_get(seekTable.data(), f), fb = entry_get(seekTable.data(), f + 1);
      u64 begin = f * header.frameSize;
      d.srcOff = fa - a;
      d.srcLen = (u32)(fb - fa);
      d.dstOff = begin - first * header.frameSize;
      d.dstCap = (u32)std::min<u64>(header.frameSize, header.uncompressedSize - begin);
      d.exact = 1;
      d.pad = 0;
    }
    raise(host_decode_frames(g, input, compressedSize, frames.data(), frames.size(), header.frameSize, r, size, output.data), g);
  }

  void Decompressor::Decompress(size_t offset, size_t size, Buffer& output) {
    output.resize(size);
    Decompress(offset, size, BufferView(output));
  }

  Buffer Decompressor::Decompress(size_t offset, size_t size) {
    Buffer buffer;
    Decompress(offset, size, buffer);
    return buffer;
  }

  // ------------------------------------------------------------------ FullDecompressor (source/zra.cpp:426-436)
  FullDecompressor::FullDecompressor(const std::function<void(size_t offset, size_t size, void* buffer)>& readFunction)
      : ctx(std::make_shared<ZDCtx>()),
        readFunction(readFunction),
        header(readFunction),
        seekTable(header.GetSeekTable()),
        entry(reinterpret_cast<Entry*>(seekTable.data())) {}

  size_t FullDecompressor::Decompress(const BufferView& output) {
    if (output.size < header.frameSize) throw Exception(StatusCode::OutputBufferTooSmall);
    if (!header.frameSize) throw Exception(StatusCode::HeaderInvalid);
    size_t entries = seekTable.size() / kEntrySize;
    size_t cur = static_cast<size_t>(reinterpret_cast<u8*>(entry) - seekTable.data()) / kEntrySize;
    size_t lastIdx = std::min(entries ? entries - 1 : 0, cur + output.size / header.frameSize);
    u64 a = entry_get(seekTable.data(), cur), b = entry_get(seekTable.data(), lastIdx);
    auto check_range = [&](size_t c0, size_t c1, u64 lo, u64 hi) {  // untrusted seek table: see Decompressor::Decompress
      if (hi < lo) throw Exception(StatusCode::ZStdError, 72);
      for (size_t f = c0; f < c1; f++) {
        const u64 fa = entry_get(seekTable.data(), f), fb = entry_get(seekTable.data(), f + 1);
        if (fa < lo || fb < fa || fb > hi || fb - fa > 0xFFFFFFFFull) throw Exception(StatusCode::ZStdError, 72);
      }
    };
    check_range(cur, lastIdx, a, b);
    // One read callback per call, like the reference (zra.cpp:431-433), into page-locked staging (see Decompressor) —
    // but the callback that fills a call's bytes usually ran during the PREVIOUS call: the reference's loop is callback,
    // decode, callback, decode on one thread, and with the decode on the GPU the host would sit idle in one half and the
    // GPU in the other. So while this call's uploads / kernels / downloads run, the NEXT call's range (same output
    // capacity assumed: the class is a sequential reader) is read ahead through the same callback, on the caller's
    // thread, into the second staging buffer. A call whose range is not the one read ahead reads for itself.
    // ZRA_B200_NO_READAHEAD=1 restores the strict callback-per-call order.
    GpuContext* g = gpu();
    const size_t compressedSize = b - a;
    static const bool readAhead = getenv("ZRA_B200_NO_READAHEAD") == nullptr;
    u8* input;
    if (ctx->aheadValid && ctx->aheadCur == cur && ctx->aheadLast == lastIdx && ctx->aheadA == a && ctx->aheadB == b &&
        g->stage_gen(ctx->aheadBuf) == ctx->aheadGen) {
      input = g->pinned_stage(compressedSize, ctx->aheadBuf);
      ctx->nextBuf = ctx->aheadBuf ^ 1;
    } else {
      input = g->pinned_stage(compressedSize, ctx->nextBuf);
      if (!input) { cache.resize(compressedSize); input = cache.data(); }
      readFunction(header.size + a, compressedSize, input);
      ctx->nextBuf ^= 1;
    }
    ctx->aheadValid = false;
    entry = reinterpret_cast<Entry*>(seekTable.data() + kEntrySize * lastIdx);
    if (lastIdx == cur) return 0;

    std::vector<HostFrame> frames(lastIdx - cur);
    size_t total = 0;
    for (size_t f = cur; f < lastIdx; f++) {
      HostFrame& d = frames[f - cur];
      u64 fa = entry_get(seekTable.data(), f), fb = entry_get(seekTable.data(), f + 1);
      u64 begin = (u64)f * header.frameSize;
      d.srcOff = fa - a;
      d.srcLen = (u32)(fb - fa);
      d.dstOff = begin - (u64)cur * header.frameSize;
      d.dstCap = (u32)std::min<u64>(header.frameSize, header.uncompressedSize - begin);
      d.exact = 1;
      d.pad = 0;
      total += d.dstCap;
    }
    // the next call's range, if there is one
    const size_t nCur = lastIdx, nLast = std::min(entries ? entries - 1 : 0, nCur + output.size / header.frameSize);
    std::function<void()> ahead = [&]() {
      // runs while this call's GPU work is in flight: nothing may unwind from here (a bad range or a throwing callback
      // simply leaves no read-ahead; the next call reads for itself and reports it)
      try {
        const u64 na = entry_get(seekTable.data(), nCur), nb = entry_get(seekTable.data(), nLast);
        check_range(nCur, nLast, na, nb);
        u8* buf = g->pinned_stage(nb - na, ctx->nextBuf);
        if (!buf) return;
        readFunction(header.size + na, nb - na, buf);
        ctx->aheadCur = nCur; ctx->aheadLast = nLast; ctx->aheadA = na; ctx->aheadB = nb; ctx->aheadBuf = ctx->nextBuf;
        ctx->aheadGen = g->stage_gen(ctx->nextBuf);
        ctx->aheadValid = true;
      } catch (...) {
        ctx->aheadValid = false;
      }
    };
    const bool doAhead = readAhead && nLast > nCur && input != cache.data();
    raise(host_decode_frames(g, input, compressedSize, frames.data(), frames.size(), header.frameSize, 0, total, output.data,
                             doAhead ? &ahead : nullptr), g);
    return total;
  }
}  // namespace zra

// ====================================================================== C shim (source/zra.cpp:439-625)
namespace {
  ZraStatus make_status(ZraStatusCode zra, int zstd = 0) { return ZraStatus{zra, static_cast<int8_t>(zstd)}; }
  ZraStatus make_status(const zra::Exception& e) { return make_status(static_cast<ZraStatusCode>(e.code), e.zstdCode); }
  const ZraStatus kOk{Success, 0};
  // nothing may unwind through the C ABI: a host failure (std::bad_alloc, std::length_error from a caller-supplied size)
  // is reported as a zstd memory_allocation error (ZSTD_error_memory_allocation = 64)
  const ZraStatus kHostFailure{ZStdError, 64};
}  // namespace

extern "C" {

uint16_t ZraGetVersion() { return zra::GetVersion(); }

const char* ZraGetErrorString(ZraStatus status) {
  return zra::Exception(static_cast<zra::StatusCode>(status.zra), status.zstd).what();
}

ZraStatus ZraCreateHeader(ZraHeader** header, void (*readFunction)(size_t, size_t, void*)) {
  try {
    *header = reinterpret_cast<ZraHeader*>(new zra::Header(readFunction));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

ZraStatus ZraCreateHeader2(ZraHeader** header, void* buffer, size_t size) {
  try {
    *header = reinterpret_cast<ZraHeader*>(new zra::Header(zra::BufferView(buffer, size)));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

void ZraDeleteHeader(ZraHeader* header) { delete reinterpret_cast<zra::Header*>(header); }
size_t ZraGetVersionWithHeader(ZraHeader* header) { return reinterpret_cast<zra::Header*>(header)->version; }
size_t ZraGetHeaderSizeWithHeader(ZraHeader* header) { return reinterpret_cast<zra::Header*>(header)->size; }
size_t ZraGetUncompressedSizeWithHeader(ZraHeader* header) { return reinterpret_cast<zra::Header*>(header)->uncompressedSize; }
size_t ZraGetFrameSizeWithHeader(ZraHeader* header) { return reinterpret_cast<zra::Header*>(header)->frameSize; }
size_t ZraGetMetadataSize(ZraHeader* header) { return reinterpret_cast<zra::Header*>(header)->metaSize; }

void ZraGetMetadata(ZraHeader* header, void* buffer) {
  auto* h = reinterpret_cast<zra::Header*>(header);
  h->GetMetadata(zra::BufferView(buffer, h->metaSize));
}

size_t ZraGetCompressedOutputBufferSize(size_t inputSize, size_t frameSize) {
  return zra::GetOutputBufferSize(inputSize, static_cast<uint32_t>(frameSize));
}

ZraStatus ZraCompressBuffer(void* inputBuffer, size_t inputSize, void* outputBuffer, size_t* outputSize, int8_t compressionLevel,
                            uint32_t frameSize, bool checksum, void* metaBuffer, size_t metaSize) {
  try {
    *outputSize = zra::CompressBuffer(zra::BufferView(inputBuffer, inputSize),
                                      zra::BufferView(outputBuffer, zra::GetOutputBufferSize(inputSize, frameSize)), compressionLevel,
                                      frameSize, checksum, zra::BufferView(metaBuffer, metaSize));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

ZraStatus ZraDecompressBuffer(void* inputBuffer, size_t inputSize, void* outputBuffer) {
  try {
    size_t cap = inputSize >= 26 ? zrab::get_le(static_cast<uint8_t*>(inputBuffer) + 18, 8) : 0;
    zra::DecompressBuffer(zra::BufferView(inputBuffer, inputSize), zra::BufferView(outputBuffer, cap));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

ZraStatus ZraDecompressRA(void* inputBuffer, size_t inputSize, void* outputBuffer, size_t offset, size_t size) {
  try {
    zra::DecompressRA(zra::BufferView(inputBuffer, inputSize), zra::BufferView(outputBuffer, size), offset, size);
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

ZraStatus ZraCreateCompressor(ZraCompressor** compressor, size_t size, int8_t compressionLevel, uint32_t frameSize, bool checksum,
                              void* metaBuffer, size_t metaSize) {
  try {
    *compressor = reinterpret_cast<ZraCompressor*>(
        new zra::Compressor(size, compressionLevel, frameSize, checksum, zra::BufferView(metaBuffer, metaSize)));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

void ZraDeleteCompressor(ZraCompressor* compressor) { delete reinterpret_cast<zra::Compressor*>(compressor); }

size_t ZraGetOutputBufferSizeWithCompressor(ZraCompressor* compressor, size_t inputSize) {
  return reinterpret_cast<zra::Compressor*>(compressor)->GetOutputBufferSize(inputSize);
}

ZraStatus ZraCompressWithCompressor(ZraCompressor* compressor, void* inputBuffer, size_t inputSize, void* outputBuffer,
                                    size_t* outputSize) {
  try {
    auto* c = reinterpret_cast<zra::Compressor*>(compressor);
    *outputSize = c->Compress(zra::BufferView(inputBuffer, inputSize), zra::BufferView(outputBuffer, c->GetOutputBufferSize(inputSize)));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

size_t ZraGetHeaderSizeWithCompressor(ZraCompressor* compressor) { return reinterpret_cast<zra::Compressor*>(compressor)->GetHeaderSize(); }

ZraStatus ZraGetHeaderWithCompressor(ZraCompressor* compressor, void* outputBuffer) {
  try {
    const auto& header = reinterpret_cast<zra::Compressor*>(compressor)->GetHeader();
    std::memcpy(outputBuffer, header.data(), header.size());
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

ZraStatus ZraCreateDecompressor(ZraDecompressor** decompressor, void (*readFunction)(size_t, size_t, void*), size_t maxCacheSize) {
  try {
    *decompressor = reinterpret_cast<ZraDecompressor*>(new zra::Decompressor(readFunction, maxCacheSize));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

void ZraDeleteDecompressor(ZraDecompressor* decompressor) { delete reinterpret_cast<zra::Decompressor*>(decompressor); }

ZraHeader* ZraGetHeaderWithDecompressor(ZraDecompressor* decompressor) {
  return reinterpret_cast<ZraHeader*>(&reinterpret_cast<zra::Decompressor*>(decompressor)->header);
}

ZraStatus ZraDecompressWithDecompressor(ZraDecompressor* decompressor, size_t offset, size_t size, void* outputBuffer) {
  try {
    reinterpret_cast<zra::Decompressor*>(decompressor)->Decompress(offset, size, zra::BufferView(outputBuffer, size));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

// maxCacheSize is accepted and ignored, exactly like the reference (source/zra.cpp:602-604)
ZraStatus ZraCreateFullDecompressor(ZraFullDecompressor** decompressor, void (*readFunction)(size_t, size_t, void*), size_t) {
  try {
    *decompressor = reinterpret_cast<ZraFullDecompressor*>(new zra::FullDecompressor(readFunction));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

void ZraDeleteFullDecompressor(ZraFullDecompressor* decompressor) { delete reinterpret_cast<zra::FullDecompressor*>(decompressor); }

ZraHeader* ZraGetHeaderWithFullDecompressor(ZraFullDecompressor* decompressor) {
  return reinterpret_cast<ZraHeader*>(&reinterpret_cast<zra::FullDecompressor*>(decompressor)->header);
}

ZraStatus ZraDecompressWithFullDecompressor(ZraFullDecompressor* decompressor, void* outputBuffer, size_t outputCapacity,
                                            size_t* outputSize) {
  try {
    *outputSize = reinterpret_cast<zra::FullDecompressor*>(decompressor)->Decompress(zra::BufferView(outputBuffer, outputCapacity));
    return kOk;
  } catch (const zra::Exception& e) { return make_status(e); } catch (const std::exception&) { return kHostFailure; }
}

}  // extern "C"

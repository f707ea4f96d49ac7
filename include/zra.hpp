// zra.hpp — C++17 interface of zra-b200 (B200-native ZRA).
//
// Drop-in contract: namespace, type names, member order, default arguments and exception
// behaviour are those of the reference's C++ interface (/root/reference/include/zra.hpp:21-324,
// implemented at source/zra.cpp:18-437), so sources written against the reference compile and
// behave the same against this library. What is different is underneath: ZCCtx / ZDCtx, which
// the reference leaves as incomplete types wrapping a zstd context, are this library's GPU
// contexts (CUDA stream, device scratch, pinned staging); every frame is encoded / decoded by
// sm_100a kernels. All pointers in this interface are host pointers.
#pragma once

#include <cstdint>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#if defined(ZRA_EXPORT_HEADER)
#include "zra_export.h"
#elif !defined(ZRA_EXPORT)
#if defined(_WIN32)
#define ZRA_EXPORT __declspec(dllimport)
#else
#define ZRA_EXPORT __attribute__((visibility("default")))
#endif
#endif

namespace zra {
  using u8 = std::uint8_t;
  using u16 = std::uint16_t;
  using u32 = std::uint32_t;
  using u64 = std::uint64_t;
  using i8 = std::int8_t;
  using i16 = std::int16_t;
  using i32 = std::int32_t;
  using i64 = std::int64_t;

  using Buffer = std::vector<u8>;  // owning byte buffer

  // Non-owning (pointer, length) pair.
  struct BufferView {
    u8* data{nullptr};
    size_t size{};

    BufferView() = default;
    constexpr inline BufferView(void* data, size_t size) : data(static_cast<u8*>(data)), size(size) {}
    inline BufferView(const Buffer& buffer) : data(const_cast<u8*>(buffer.data())), size(buffer.size()) {}
  };

  // Same numbering as ZraStatusCode in zra.h.
  enum class StatusCode {
    Success,
    ZStdError,
    ZraVersionLow,
    HeaderInvalid,
    HeaderIncomplete,
    OutOfBoundsAccess,
    OutputBufferTooSmall,
    CompressedSizeTooLarge,
    InputFrameSizeMismatch,
  };

  // Thrown by every function below on failure.
  // reference: include/zra.hpp:70-88, source/zra.cpp:48-82
  struct ZRA_EXPORT Exception : std::exception {
    StatusCode code;  // what failed
    int zstdCode;     // ZSTD_ErrorCode when code == ZStdError

    Exception(StatusCode code, i32 zstdCode = {});
    static const char* GetExceptionString(StatusCode code);
    const char* what() const noexcept override;
  };

  // Highest archive format version supported (1).
  // reference: include/zra.hpp:91, source/zra.cpp:84-87
  ZRA_EXPORT u16 GetVersion();

  // Parsed archive header. Keeps the read callback to fetch metadata / seek table on demand.
  // reference: include/zra.hpp:96-131, source/zra.cpp:141-187
  class ZRA_EXPORT Header {
   private:
    std::function<void(size_t, size_t, void*)> readFunction;

   public:
    u16 version;           // format version of the archive
    u32 size;              // bytes before the first frame (fixed part + metadata + seek table)
    u64 uncompressedSize;  // length of the original data
    u32 frameSize;         // uncompressed bytes per frame (the last frame may be shorter)
    u32 metaOffset;        // where the metadata section starts
    u32 metaSize;          // its length
    u32 seekTableOffset;   // where the seek table starts
    u32 seekTableSize;     // its length in bytes (5 per entry)

    // readFunction(offset, size, buffer) must copy `size` archive bytes at `offset` into `buffer`.
    Header(const std::function<void(size_t offset, size_t size, void* buffer)>& readFunction);
    // For an archive that is entirely in memory.
    Header(const BufferView& buffer);

    void GetMetadata(const BufferView& buffer) const;
    Buffer GetMetadata() const;
    Buffer GetSeekTable() const;
  };

  // Worst-case archive size for inputSize bytes in frameSize frames with metaSize metadata bytes.
  // reference: include/zra.hpp:139, source/zra.cpp:189-192
  ZRA_EXPORT size_t GetOutputBufferSize(size_t inputSize, u32 frameSize, u32 metaSize = 0);

  // Whole-buffer compression into `output` (capacity >= GetOutputBufferSize); returns the archive size.
  // reference: include/zra.hpp:151, source/zra.cpp:194-234
  ZRA_EXPORT size_t CompressBuffer(const BufferView& input, const BufferView& output, i8 compressionLevel = 0,
                                   u32 frameSize = 16384, bool checksum = true, const BufferView& meta = {});
  // Same, returning a right-sized Buffer.
  // reference: include/zra.hpp:162, source/zra.cpp:236-241
  ZRA_EXPORT Buffer CompressBuffer(const BufferView& buffer, i8 compressionLevel = 0, u32 frameSize = 16384,
                                   bool checksum = true, const BufferView& meta = {});

  // Whole-archive decompression; `output` must hold Header::uncompressedSize bytes.
  // reference: include/zra.hpp:169,176, source/zra.cpp:243-256
  ZRA_EXPORT void DecompressBuffer(const BufferView& input, const BufferView& output);
  ZRA_EXPORT Buffer DecompressBuffer(const BufferView& buffer);

  // Random access: `size` bytes of the original data starting at `offset`.
  // reference: include/zra.hpp:185,194, source/zra.cpp:258-302
  ZRA_EXPORT void DecompressRA(const BufferView& input, const BufferView& output, size_t offset, size_t size);
  ZRA_EXPORT Buffer DecompressRA(const BufferView& buffer, size_t offset, size_t size);

  class ZCCtx;   // GPU compression context (opaque)
  struct Entry;  // 5-byte seek-table slot (opaque)

  // Streaming compression: feed the input in frame-aligned chunks, write the header last.
  // reference: include/zra.hpp:202-254, source/zra.cpp:304-365
  class ZRA_EXPORT Compressor {
   private:
    std::shared_ptr<ZCCtx> ctx;
    u32 frameSize;
    u32 tableSize;
    Buffer header;
    Entry* entry;
    size_t outputOffset{};

   public:
    // `size` is the exact length of the whole stream.
    Compressor(size_t size, i8 compressionLevel = 0, u32 frameSize = 16384, bool checksum = true, const BufferView& meta = {});

    // Worst-case output of one Compress() call on inputSize bytes (not the archive bound).
    size_t GetOutputBufferSize(size_t inputSize) const;

    // Compresses the next chunk; returns the compressed size written to `output`.
    size_t Compress(const BufferView& input, const BufferView& output);
    void Compress(const BufferView& input, Buffer& output);

    // Complete header; throws HeaderIncomplete until the final chunk has been compressed.
    const Buffer& GetHeader();
    size_t GetHeaderSize();
  };

  class ZDCtx;  // GPU decompression context (opaque)

  // Streaming random access over a read callback.
  // reference: include/zra.hpp:261-299, source/zra.cpp:367-424
  class ZRA_EXPORT Decompressor {
   private:
    std::shared_ptr<ZDCtx> ctx;
    std::function<void(size_t, size_t, void*)> readFunction;

   public:
    Header header;

   private:
    Buffer seekTable;
    Buffer cache;
    size_t maxCacheSize;

   public:
    Decompressor(const std::function<void(size_t offset, size_t size, void* buffer)>& readFunction,
                 size_t maxCacheSize = 1024 * 1024 * 20);

    void Decompress(size_t offset, size_t size, const BufferView& output);
    void Decompress(size_t offset, size_t size, Buffer& output);
    Buffer Decompress(size_t offset, size_t size);
  };

  // Streaming front-to-back decompression over a read callback.
  // reference: include/zra.hpp:303-324, source/zra.cpp:426-436
  class ZRA_EXPORT FullDecompressor {
   private:
    std::shared_ptr<ZDCtx> ctx;
    std::function<void(size_t, size_t, void*)> readFunction;

   public:
    Header header;

   private:
    Buffer seekTable;
    Buffer cache;
    Entry* entry;

   public:
    FullDecompressor(const std::function<void(size_t offset, size_t size, void* buffer)>& readFunction);

    // Decodes as many whole frames as fit `output` (>= one frame); returns bytes produced, 0 at the end.
    size_t Decompress(const BufferView& output);
  };
}  // namespace zra

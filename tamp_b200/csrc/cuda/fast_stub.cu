// Placeholder dispatch for configurations without a specialised kernel (replaced file by file as the
// specialised kernels land; see fast_compress.cu / fast_decompress.cu).
#include "tb_cuda.h"
#define TB_HAVE_FAST_COMPRESS 1
#define TB_HAVE_FAST_DECOMPRESS 1
namespace tb {
#ifndef TB_HAVE_FAST_COMPRESS
bool launch_fast_compress_batch(const CompBatchConf &, const uint8_t *, const BatchArgs &, cudaStream_t) { return false; }
#endif
#ifndef TB_HAVE_FAST_DECOMPRESS
bool launch_fast_decompress_batch(const uint8_t *, const uint8_t *, int, const BatchArgs &, cudaStream_t) { return false; }
#endif
}  // namespace tb

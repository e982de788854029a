// placeholder until the ReID kernels land (next commit)
#include <stdexcept>
#include "engine.h"
namespace bmb {
struct ReidModel { int dim; };
ReidModel* reid_load(const char*) { throw std::runtime_error("ReID model support not built"); }
void reid_free(ReidModel* m) { delete m; }
int reid_feature_dim(const ReidModel* m) { return m->dim; }
int reid_forward(ReidModel*, const uint8_t*, size_t, int, int, const CropDesc*, const int*, int, float*, int, cudaStream_t) { return 0; }
const float* reid_last_input_blob(const ReidModel*) { return nullptr; }
}

// Allegro B200 pipeline instantiation for l_max = 3 (FP32-pipe path and tensor-core path)
#define ALG_PIPELINE_IMPL
#include "alg_pipeline.cuh"
namespace alg {
ALG_DEFINE_PIPELINE(3)
ALG_DEFINE_PIPELINE_TC(3)
}

// Allegro B200 pipeline instantiation for l_max = 1
#define ALG_PIPELINE_IMPL
#include "alg_pipeline.cuh"
namespace alg {
ALG_DEFINE_PIPELINE(1)
}

// Allegro B200 pipeline instantiation for l_max = 2
#define ALG_PIPELINE_IMPL
#include "alg_pipeline.cuh"
namespace alg {
ALG_DEFINE_PIPELINE(2)
}

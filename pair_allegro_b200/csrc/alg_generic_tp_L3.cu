// width-generic pipeline: tensor-product kernels for l_max = 3
#include "alg_generic_tp.cuh"
namespace alg {
ALG_DEFINE_GENERIC_TP(3)
}

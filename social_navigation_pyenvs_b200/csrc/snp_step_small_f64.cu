// The fused small-crowd step kernel (snp_step_small.inl) instantiated for double; one translation unit per arithmetic type so
// the two sets of model / mapping instantiations compile in parallel.
#define SNP_STEP_DTYPE double
#include "snp_step_small.inl"

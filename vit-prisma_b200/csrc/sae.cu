// sae.cu -- sparse-autoencoder training path (placeholder until the kernels land in this file).
#include "common.cuh"
int pb_abi_sizeof_sae(int which) { (void)which; return -1; }

// Single translation unit of libboojum_b200.so (keeps __constant__ tables and inlining in one module).
#include "capi.cu"
#include "ntt.cu"
#include "poseidon2.cu"
#include "fri.cu"
#include "elementwise.cu"
#include "gates.cu"
#include "host_transcript.cu"
#include "fri_driver.cu"

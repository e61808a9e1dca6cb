// checkpoint.h -- pool checkpoint / resume: the structure-of-arrays region pool and the loop
// counters of a DEVICE-mode integration on one GPU as a file.  The reference has no counterpart (its heap lives and dies
// inside Rule.cubature, Rule.java:51-160); the use case is examples/PlotEvalPositions.java:28-54 (what did the adaptive loop
// evaluate?) and continuing a maxEval-limited run with a larger budget or a tighter tolerance.
//
// File layout (little endian): CkptHeader, then the rows center[dim][n], halfw[dim][n], val[fdim][n], err[fdim][n],
// errmax[n] as doubles and splitdim[n] as int32, each row dense (n elements).
#ifndef CUB_CHECKPOINT_H
#define CUB_CHECKPOINT_H

#include <cstdint>
#include <cstdio>

#include "cub_internal.h"

struct CkptHeader {
    char magic[8];  // "CUBCKPT1"
    int32_t version, dim, fdim, family;
    int64_t n_regions, num_eval, iterations;
    int32_t transform[CUB_MAX_DIM];
    double shift[CUB_MAX_DIM], tmin[CUB_MAX_DIM], tmax[CUB_MAX_DIM];
    uint64_t integrand_hash;  // family, dimensions, parameters, user source
    uint64_t payload_checksum;
    uint64_t header_checksum;  // over the bytes before this field
};

uint64_t ckpt_integrand_hash(const cub_integrand* integ);
uint64_t ckpt_mix(uint64_t h, const void* data, size_t bytes);
// reads and verifies the header (not the payload); returns CUB_OK or CUB_ERR_BAD_ARGS with a message
int ckpt_read_header(FILE* f, CkptHeader* h, char* msg);
void ckpt_seal_header(CkptHeader* h);
size_t ckpt_payload_bytes(const CkptHeader& h);

#endif

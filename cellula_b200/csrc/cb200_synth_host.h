// cb200_synth_host.h -- host pieces of the synthetic-data generator shared by cb200_host.cu (host generator)
// and cb200_synth.cu (device generator): the G-sized model parameters and the density calibration are
// computed by the same host code for both, so that the two generators tile the same matrix.
#pragma once
#include <cstdint>
#include <vector>

#define CB200_SYNTH_PROBE 2048  // cells of the global matrix whose realised density calibrates the offset

void cb200_synth_params_host(int64_t G, uint64_t seed, std::vector<float> &base, std::vector<uint8_t> &wf,
                             std::vector<float> &wv, std::vector<float> &cent);
float cb200_synth_neutral_offset(const std::vector<float> &base, double nnz_per_cell);
float cb200_synth_refine_offset(float off, double nnz_per_cell, int64_t nnz_probe, int64_t n_probe);
bool cb200_synth_offset_ok(double nnz_per_cell, int64_t nnz_probe, int64_t n_probe);

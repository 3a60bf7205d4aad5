// placeholder until the ViT forward lands (same symbols, loud failure)
#include "d2r_common.cuh"
#define NOT_YET(name) { d2r::set_error(name ": not implemented yet"); return D2R_ERR_INVALID; }
extern "C" int d2r_clip_preprocess(const uint8_t*, int, int, int, int, int, int, const float*, const float*, void*, float*, void*) NOT_YET("d2r_clip_preprocess")
extern "C" int d2r_clip_load(const d2r_clip_cfg*, const float* const*, int, int, d2r_clip**) NOT_YET("d2r_clip_load")
extern "C" void d2r_clip_free(d2r_clip*) {}
extern "C" int d2r_clip_encode(d2r_clip*, const void*, int, float*, void*) NOT_YET("d2r_clip_encode")
extern "C" int d2r_score(const float*, const float*, int, int, int, float, int, float*, float*, void*) NOT_YET("d2r_score")

/*
 * features_device.h - device-side description of a feature space (features.cu) and its launchers.
 */
#ifndef FDB_FEATURES_DEVICE_H_
#define FDB_FEATURES_DEVICE_H_

#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#include "fdb_internal.h"
#include "wvm_device.h"

namespace fdb {

struct FeatCache { /* HistogramFilter::CacheEntry (HistogramFilter.hpp:73-78) */
	int index1, index2;
	float weight1, weight2;
};

struct FeatureShape {
	int dim, is_float, layer_channels, bins, cell_rows, cell_cols, use_hog_filter;
};

struct DevFeature {
	int kind, pw, ph, dim, is_float;
	int layer_channels;          /* bytes per pixel of the filtered layers; 0: the chain works on the gray layer */
	int bins, cell_rows, cell_cols, use_hog_filter;
	int block_size, concatenate, signed_and_unsigned, normalization, interpolate_cells;
	int gradient_kernel, lbp_type;
	float ehog_alpha;
	int n_layers;
	const uint8_t* lut;          /* GradientBinningFilter table [65536][layer_channels] */
	const FeatCache* row_cache;  /* [ph], [pw]: bilinear cell interpolation */
	const FeatCache* col_cache;
	const float* whi_filter;     /* [ph][pw] */
	const double* twiddle;       /* cos_w[pw], sin_w[pw], cos_h[ph], sin_h[ph] */
	int64_t layer_offset[FDB_MAX_LAYERS];   /* byte offset of each filtered layer inside the per-frame feature arena */
	int px_prefix[FDB_MAX_LAYERS + 1];      /* prefix sums of the layer pixel counts */
	uint8_t lbp_map[256];
};

int feature_shape(const fdb_feature_desc& d, int pw, int ph, FeatureShape* s);
int feature_build(const fdb_feature_desc& d, int pw, int ph, const Plan& plan, DevFeature* out, int64_t* farena_bytes,
		std::vector<void*>& owned);
int feature_configure();
size_t feature_smem_bytes(const DevFeature& f);
void launch_feature_layers(cudaStream_t st, const DevFeature& f, const uint8_t* frames, int W, int H, int n_frames,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, uint8_t* farena, int64_t farena_stride);
void launch_feature_patches(cudaStream_t st, const DevFeature& f, const uint8_t* frames, int W, int H,
		const uint8_t* arena, int64_t arena_stride, const DevLayer* layers, const uint8_t* farena, int64_t farena_stride,
		const SvmItem* items, int n_items, void* out);

} // namespace fdb
#endif

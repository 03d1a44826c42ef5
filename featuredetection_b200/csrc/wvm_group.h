/*
 * wvm_group.h - tables of the group window kernel and the multi-detector deep kernel (wvm_group.cu).
 *
 * A "group" is what several detectors of one application have in common: ffpDetectApp runs 15 landmark detectors over
 * every frame (ffpDetectApp.cpp:548-596), 12 of which scan the SAME four pyramid layers and 7 of those with the SAME
 * 24 x 24 window - HistEq64 of a window depends only on (layer, position, window size), so one equalisation serves
 * every model of the same geometry. A work item is a strip of windows of one layer image and a PACK of up to
 * GRP_MAX_PACK models evaluated on it.
 */
#ifndef FDB_WVM_GROUP_H_
#define FDB_WVM_GROUP_H_

#include <cuda_runtime.h>
#include <cstdint>

#include "wvm_device.h"

namespace fdb {

#define GRP_MAX_MODELS 16 /* detectors per launch (ffpDetectApp: 15) */
#ifndef GRP_MAX_PACK
#define GRP_MAX_PACK 4    /* models sharing one equalisation inside the window kernel (accumulators in tensor memory: 32 columns per model) */
#endif

struct GroupModel {       /* one detector's stage-1 classifier and where its results go */
	DevWvm m;
	fdb_window_score* dense;   /* [frame][windows_per_frame] or null */
	int windows_per_frame;
	int cand_cap;
	Candidate* cand;           /* positives of the launch, unordered; null: none wanted */
	int* cand_count;
	DeepQueue q;               /* survivors of the first WVM_KA filters */
};

struct GroupImage {       /* a pyramid image windows are cut from */
	int64_t offset;       /* arena offset; < 0: the input frame itself */
	int width, height, pitch;
	int tma_ok;           /* a tensor map exists (arena image) */
};

struct GroupItem {
	int image;                        /* GroupImage index */
	int begin_x, begin_y;             /* first window corner of the layer scan */
	int windows_x, windows_y;
	int ix0, iy0, cols, nsub, run;    /* the strip: window columns ix0.., nsub row runs of `run` rows from iy0 */
	int nm;                           /* models in the pack */
	int model[GRP_MAX_PACK];          /* GroupModel index */
	int first_window[GRP_MAX_PACK];   /* canonical index of the layer's first window in that detector's order */
	int pad;
};

struct GroupArgs {
	const GroupItem* items; int n_items; int n_frames;
	const GroupImage* images;
	const void* tmaps;               /* CUtensorMap[image] or null */
	const uint8_t* frames; int W, H;
	const uint8_t* arena; int64_t arena_stride;
	int* cursor;                     /* work counter (zeroed before the launch) */
	GroupModel models[GRP_MAX_MODELS];
};

struct DeepArgs {
	const GroupImage* images;
	const uint8_t* frames; int W, H;
	const uint8_t* arena; int64_t arena_stride;
	int n_models;
	GroupModel models[GRP_MAX_MODELS];
};

int group_configure_all();
int group_tc_configure_all();
/* models per pack: GRP_MAX_PACK if every model has the tcgen05 operand (DevWvm::btc), or 2 - also when FDB_WINDOW_KERNEL=mma
 * rules the tcgen05 kernel out */
int group_max_pack(bool tc_ok);
bool group_supported(int patch_w, int patch_h);
/* window kernel over args.items (all of one window size, packs of at most `pack` models); tc_ok: every model has DevWvm::btc */
void launch_wvm_group(cudaStream_t st, int patch_w, int patch_h, int pack, const GroupArgs& args, bool tc_ok);
void launch_wvm_group_mma(cudaStream_t st, int patch_w, int patch_h, int pack, const GroupArgs& args);
void launch_wvm_group_tc(cudaStream_t st, int patch_w, int patch_h, int pack, const GroupArgs& args);
/* the rest of the cascade for every queued survivor of every model of the table: one launch */
void launch_wvm_deep_group(cudaStream_t st, const DeepArgs& args);

} // namespace fdb
#endif

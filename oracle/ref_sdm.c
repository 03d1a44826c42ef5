/*
 * ref_sdm.c - C entry point over the reference's OWN vendored VLFeat HOG
 * (libSupervisedDescent/src/superviseddescent/hog.c, compiled unmodified into oracle/_ref by oracle/Makefile),
 * called exactly as VlHogDescriptorExtractor::getDescriptors does (DescriptorExtractor.hpp:185-193).
 * TEST INFRASTRUCTURE ONLY: pins fdo_vlhog_uoctti (oracle/fd_sdm.c).
 */
#include "superviseddescent/hog.h"

int ref_vlhog_uoctti(const float* image, int width, int height, int cell_size, int num_orientations, float* features,
		int* hw_out, int* hh_out) {
	VlHog* hog = vl_hog_new(VlHogVariantUoctti, (vl_size)num_orientations, 0);
	vl_hog_put_image(hog, image, (vl_size)width, (vl_size)height, 1, (vl_size)cell_size);
	const int ww = (int)vl_hog_get_width(hog), hh = (int)vl_hog_get_height(hog), dd = (int)vl_hog_get_dimension(hog);
	vl_hog_extract(hog, features);
	vl_hog_delete(hog);
	if (hw_out) *hw_out = ww;
	if (hh_out) *hh_out = hh;
	return dd;
}

// Host-side writers for the render path's outputs (csrc/image_out.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ptb
{

// rgb: height x width x 3, row 0 = top.  Written as 8-bit RGBA (alpha 255) like Main/window.cpp:712-740.
bool write_png_rgb8(const std::string& path, const uint8_t* rgb, int width, int height, std::string& err);
// rgb * scale as a little-endian colour PFM (rows bottom to top)
bool write_pfm_rgb(const std::string& path, const float* rgb, int width, int height, float scale, std::string& err);

// the reference's `image` state that must survive a restart: size + pass_counter (Core/image.h:10-23), plus the
// camera the passes were rendered with and the depth limit that fixes the per-pass clamp
struct CheckpointHeader
{
	int32_t width, height, pass_counter, max_depth;
	float camera[16];
};

bool write_checkpoint(const std::string& path, const CheckpointHeader& h, const float* sum_rgb, std::string& err);
bool read_checkpoint(const std::string& path, CheckpointHeader& h, std::vector<float>& sum_rgb, std::string& err);

} // namespace ptb

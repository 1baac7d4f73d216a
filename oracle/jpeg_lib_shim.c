/* TEST INFRASTRUCTURE — never linked by the product.
 *
 * Drives an installed libjpeg (the libjpeg-turbo that ships inside Pillow's wheel: pillow.libs/libjpeg-*.so.62) through its
 * public C API with the two decompression parameters FreeImage sets for the reference's loads
 * (Others/image_loader.cpp:45 passes flags 0 = JPEG_DEFAULT = JPEG_FAST, lib/free_image/FreeImage.h:693-695; FreeImage's JPEG
 * plugin then selects dct_method = JDCT_IFAST and do_fancy_upsampling = FALSE).  PIL itself cannot select them, so the golden
 * vectors of csrc/jpeg_decode.cpp's "fast" mode come from here (tests/golden/make_jpeg_golden.py).
 *
 * No jpeglib.h in this image: the prefix of jpeg_decompress_struct up to do_block_smoothing and the output_* fields are laid out
 * below as the libjpeg ABI (v6b .. v9, libjpeg-turbo) defines them on LP64; jpeg_CreateDecompress itself checks the struct SIZE
 * (found by search) and the caller checks image_width / output_width / output_components against the file header, so a
 * mismatching layout fails instead of silently decoding with other parameters.
 *
 *   cc -O2 -fPIC -shared -o _build/libjpegshim.so jpeg_lib_shim.c -ldl
 */
#include <dlfcn.h>
#include <setjmp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
	void* err; void* mem; void* progress; void* client_data;
	int is_decompressor; int global_state;
	void* src;
	unsigned image_width, image_height;
	int num_components; int jpeg_color_space; int out_color_space;
	unsigned scale_num, scale_denom;
	double output_gamma;
	int buffered_image, raw_data_out;
	int dct_method;            /* 0 islow, 1 ifast, 2 float */
	int do_fancy_upsampling, do_block_smoothing;
	int quantize_colors, dither_mode, two_pass_quantize, desired_number_of_colors;
	int enable_1pass_quant, enable_external_quant, enable_2pass_quant;
	unsigned output_width, output_height;
	int out_color_components, output_components, rec_outbuf_height;
	int actual_number_of_colors;
	void* colormap;
	unsigned output_scanline;
	char rest[4096];
} dinfo;

static jmp_buf g_jmp;
static void on_error(void* cinfo) { (void)cinfo; longjmp(g_jmp, 1); }

/* returns 0 on success; out_rgb may be NULL to query the size */
int jpegshim_decode(const char* libpath, int abi_version, const uint8_t* data, size_t n, int dct_method, int fancy,
					int* width, int* height, uint8_t* out_rgb)
{
	void* lib = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
	if (!lib) return 10;
	void* (*std_error)(void*) = (void* (*)(void*))dlsym(lib, "jpeg_std_error");
	void (*create)(void*, int, size_t) = (void (*)(void*, int, size_t))dlsym(lib, "jpeg_CreateDecompress");
	void (*mem_src)(void*, const unsigned char*, unsigned long) = (void (*)(void*, const unsigned char*, unsigned long))dlsym(lib, "jpeg_mem_src");
	int (*read_header)(void*, int) = (int (*)(void*, int))dlsym(lib, "jpeg_read_header");
	int (*start)(void*) = (int (*)(void*))dlsym(lib, "jpeg_start_decompress");
	unsigned (*read_lines)(void*, uint8_t**, unsigned) = (unsigned (*)(void*, uint8_t**, unsigned))dlsym(lib, "jpeg_read_scanlines");
	int (*finish)(void*) = (int (*)(void*))dlsym(lib, "jpeg_finish_decompress");
	void (*destroy)(void*) = (void (*)(void*))dlsym(lib, "jpeg_destroy_decompress");
	if (!std_error || !create || !mem_src || !read_header || !start || !read_lines || !finish || !destroy) return 11;

	static char err_storage[1024];
	static dinfo ci;
	volatile int created = 0;
	for (size_t size = 400; size <= 1024 && !created; size += 8)
	{
		memset(&ci, 0, sizeof(ci));
		memset(err_storage, 0, sizeof(err_storage));
		ci.err = std_error(err_storage);
		*(void (**)(void*))err_storage = on_error;          /* jpeg_error_mgr::error_exit is the first member */
		if (setjmp(g_jmp) == 0) { create(&ci, abi_version, size); created = 1; }
	}
	if (!created) return 12;
	if (setjmp(g_jmp) != 0) { destroy(&ci); return 13; }
	mem_src(&ci, data, (unsigned long)n);
	if (read_header(&ci, 1) != 1) { destroy(&ci); return 14; }
	*width = (int)ci.image_width; *height = (int)ci.image_height;
	if (!out_rgb) { destroy(&ci); return 0; }
	ci.dct_method = dct_method;
	ci.do_fancy_upsampling = fancy;
	ci.out_color_space = 2;      /* JCS_RGB */
	start(&ci);
	if (ci.output_width != ci.image_width || ci.output_height != ci.image_height || ci.output_components != 3) { destroy(&ci); return 15; }
	for (unsigned y = 0; y < ci.output_height; y++)
	{
		uint8_t* row = out_rgb + (size_t)y * ci.output_width * 3;
		if (ci.output_scanline != y || read_lines(&ci, &row, 1) != 1) { destroy(&ci); return 16; }
	}
	finish(&ci);
	destroy(&ci);
	return 0;
}

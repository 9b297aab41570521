/* The C ABI from plain C (C99): what a cgo / JNI / N-API binding does, written out.  Loads an OBJ model, makes it
 * resident, renders one frame exactly like Viewer/Scene.cpp:32-65 + Viewer/Main.cpp:50-69 (BeginFrame, ClearFrameBuffer,
 * one DrawIndexed per mesh, EndFrame) and reads the tiles back.
 *   usage: abi_example <model.obj> <out.bin>
 * tests/test_gpu_obj.py runs it on the GPU box; tests/test_abi.py only checks that it compiles as C and links. */
#include <softrast_b200.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call, ctx)                                                                                   \
	do                                                                                                     \
	{                                                                                                      \
		int rc_ = (call);                                                                                  \
		if (rc_ != SRB_OK)                                                                                 \
		{                                                                                                  \
			fprintf(stderr, "%s failed (%d): %s %s\n", #call, rc_, (ctx) ? srb_last_error(ctx) : "", srb_model_last_error()); \
			return 1;                                                                                      \
		}                                                                                                  \
	} while (0)

int main(int argc, char** argv)
{
	const uint32_t W = 448, H = 256;
	srb_model* model = NULL;
	srb_context* ctx = NULL;
	srb_resident_model* resident = NULL;
	srb_handle fb = 0;
	srb_draw_desc draws[64];
	uint32_t n = 0, i, tiles_x = 0, tiles_y = 0, tiles;
	float mvp[16];
	float f, range;
	void *colour, *depth;
	FILE* out;

	if (argc < 3) return 2;
	CHECK(srb_model_load(argv[1], SRB_OBJ_NO_CACHE_WRITE, &model), NULL);
	CHECK(srb_create(0, SRB_FLAG_NONE, &ctx), ctx);
	CHECK(srb_framebuffer_create(ctx, W, H, &fb), ctx);
	CHECK(srb_model_make_resident(ctx, model, &resident), ctx);

	/* kt::Mat4::PerspectiveLH_ZO(85 deg, W/H, near = 10000, far = 0.1): reverse Z as in Viewer/Scene.cpp:16-29 */
	memset(mvp, 0, sizeof(mvp));
	f = tanf(1.57079632679f - 85.0f * 0.01745329252f * 0.5f);
	range = 0.1f / (0.1f - 10000.0f);
	mvp[0] = f / ((float)W / (float)H);
	mvp[5] = f;
	mvp[10] = range;
	mvp[11] = 1.0f;
	mvp[14] = -range * 10000.0f;

	CHECK(srb_resident_model_draws(resident, fb, mvp, SRB_SHADER_UNLIT_DIFFUSE, draws, 64, &n), ctx);
	CHECK(srb_begin_frame(ctx), ctx);
	CHECK(srb_clear(ctx, fb, 0, 1, 1), ctx);
	for (i = 0; i < n; ++i) CHECK(srb_draw_indexed(ctx, &draws[i]), ctx);
	CHECK(srb_end_frame(ctx), ctx);

	CHECK(srb_framebuffer_info(ctx, fb, NULL, NULL, &tiles_x, &tiles_y), ctx);
	tiles = tiles_x * tiles_y;
	colour = malloc((size_t)tiles * SRB_COLOUR_TILE_BYTES);
	depth = malloc((size_t)tiles * 16384u);
	CHECK(srb_read_tiles(ctx, fb, colour, depth, 16384), ctx);

	out = fopen(argv[2], "wb");
	if (!out) return 4;
	{
		uint32_t hdr[4];
		hdr[0] = W, hdr[1] = H, hdr[2] = tiles, hdr[3] = n;
		fwrite(hdr, sizeof(hdr), 1, out);
		fwrite(mvp, sizeof(mvp), 1, out);
		fwrite(colour, SRB_COLOUR_TILE_BYTES, tiles, out);
		fwrite(depth, 16384u, tiles, out);
	}
	fclose(out);
	printf("draws %u tiles %u launches %llu\n", n, tiles, (unsigned long long)srb_launch_count(ctx));
	free(colour);
	free(depth);
	srb_resident_model_free(resident);
	srb_destroy(ctx);
	srb_model_free(model);
	return 0;
}

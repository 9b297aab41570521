// Runs the host-side loaders (srb_model_load incl. its cache path, srb_image_load_rgba8) over the files named on the command
// line; tests/test_obj.py builds it together with srb_model.cpp / srb_host.cpp under -fsanitize=address,undefined and feeds it
// valid, fuzzed and corrupted OBJ / PNG / TGA / .bin files.  The device-side entry points the loader links against are stubbed.
#include <softrast_b200.h>
#include <stdio.h>
#include <string.h>
extern "C" {
// stubs for the device-side entry points srb_model.cpp links against (never called here)
int srb_buffer_create(srb_context*, const void*, uint64_t, srb_handle*) { return 1; }
int srb_buffer_destroy(srb_context*, srb_handle) { return 1; }
int srb_texture_create(srb_context*, const uint8_t*, uint64_t, const uint32_t*, uint32_t, uint32_t, uint32_t, srb_handle*) { return 1; }
int srb_texture_destroy(srb_context*, srb_handle) { return 1; }
const char* srb_last_error(srb_context*) { return ""; }
int srb_texture_create_rgba8(srb_context*, const uint8_t*, uint32_t, uint32_t, int, srb_handle*) { return 1; }
int srb_texture_read(srb_context*, srb_handle, uint8_t*, uint64_t, uint64_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*) { return 1; }
}
int main(int argc, char** argv)
{
	int ok = 0, bad = 0;
	for (int i = 1; i < argc; ++i)
	{
		size_t n = strlen(argv[i]);
		if (n > 4 && !strcmp(argv[i] + n - 4, ".obj"))
		{
			srb_model* m = nullptr;
			if (srb_model_load(argv[i], SRB_OBJ_NO_CACHE_WRITE, &m) == SRB_OK) { ++ok; srb_model_free(m); } else ++bad;
		}
		else if (n > 4 && !strcmp(argv[i] + n - 4, ".bin"))
		{
			// argv is "<x>.obj.bin": load through the cache path of "<x>.obj"
			char path[4096];
			snprintf(path, sizeof(path), "%.*s", (int)(n - 4), argv[i]);
			srb_model* m = nullptr;
			if (srb_model_load(path, SRB_OBJ_NO_CACHE_WRITE, &m) == SRB_OK) { ++ok; srb_model_free(m); } else ++bad;
		}
		else
		{
			uint8_t* px = nullptr; uint32_t w = 0, h = 0;
			if (srb_image_load_rgba8(argv[i], &px, &w, &h) == SRB_OK) { ++ok; srb_image_free(px); } else ++bad;
		}
	}
	printf("ok %d rejected %d\n", ok, bad);
	return 0;
}

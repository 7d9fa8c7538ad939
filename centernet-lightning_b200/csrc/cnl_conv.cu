// placeholder until the tcgen05 engine lands (next commit)
#include "cnl_common.h"
using namespace cnl;
extern "C" {
int cnl_engine_create(cnl_engine**, const cnl_buffer_desc*, int, const cnl_conv_desc*, int, int, int, int, int, int) { return fail(CNL_ERR_UNSUPPORTED, "engine not built"); }
void cnl_engine_destroy(cnl_engine*) {}
size_t cnl_engine_arena_bytes(const cnl_engine*) { return 0; }
size_t cnl_engine_buffer_offset(const cnl_engine*, int) { return 0; }
int cnl_engine_upload(cnl_engine*, void*, void*) { return fail(CNL_ERR_UNSUPPORTED, "engine not built"); }
int cnl_engine_forward(cnl_engine*, void*, const float*, int, int, void*, int*) { return fail(CNL_ERR_UNSUPPORTED, "engine not built"); }
int cnl_engine_read_buffer(cnl_engine*, void*, int, float*, void*) { return fail(CNL_ERR_UNSUPPORTED, "engine not built"); }
int cnl_engine_write_buffer(cnl_engine*, void*, int, const float*, void*) { return fail(CNL_ERR_UNSUPPORTED, "engine not built"); }
}

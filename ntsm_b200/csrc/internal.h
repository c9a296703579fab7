// internal.h -- helpers shared between the library's translation units (not exported in the header)
#pragma once
#include <stdint.h>

#include "../../include/ntsm_b200.h"

uint64_t ntsm_ctx_max_counts(const ntsm_ctx *c);
uint64_t ntsm_ctx_batch_bases(const ntsm_ctx *c);
void ntsm_set_thread_error(const char *text);
// exact -m stop after the batch just completed on ctx (ctx.cu); hits_elsewhere = hits on the other GPUs
int ntsm_trim_to_cap(ntsm_ctx *c, uint64_t hits_elsewhere, uint64_t cap);

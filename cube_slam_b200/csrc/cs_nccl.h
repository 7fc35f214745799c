/* cs_nccl.h -- NCCL is bound at run time (dlopen) so the library loads on hosts without it. */
#ifndef CS_NCCL_H
#define CS_NCCL_H
struct cs_ctx;
void cs_nccl_teardown(cs_ctx *c);
#endif

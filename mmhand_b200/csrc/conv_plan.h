// Internal: the tensor-core kernels behind mmh_conv_plan_* / mmh_wgrad_plan_* (plans.cu).
#pragma once
#include "../../include/mmhand_sm100.h"

struct MmhConv2;   // tc_conv2.cu: activation windows in shared memory, optional cta_group::2 pairs
int mmh_conv2_create(const MmhConvDesc* d, MmhConv2** out_plan);
void mmh_conv2_destroy(MmhConv2* plan);
int mmh_conv2_run(const MmhConv2* plan, void* stream);
int mmh_conv2_run_key(const MmhConv2* plan, uint32_t drop_key, void* stream);   // per-launch dropout key (bs_*)

struct MmhWgrad2;  // tc_wgrad2.cu: windows shared by the taps, pairs, taps-on-M mode for the stems
int mmh_wgrad2_create(const MmhWgradDesc* d, MmhWgrad2** out_plan);
void mmh_wgrad2_destroy(MmhWgrad2* plan);
int mmh_wgrad2_run(const MmhWgrad2* plan, void* stream);

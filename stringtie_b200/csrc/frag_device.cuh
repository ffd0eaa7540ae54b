// frag_device.cuh — row a11 (placeholder while a10 is being brought up)
#pragma once
#include "asm_device.cuh"
struct BundleGraphs { int dummy; };

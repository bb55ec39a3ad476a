// stack_common.cuh -- argument block shared by the master-combine kernels
#pragma once

#define STACK_MAX 64

struct StackArgs {
    const float *frames[STACK_MAX];     // device pointers, kernel parameter space
    float scale[STACK_MAX];             // divisor per frame (0 = none)
};

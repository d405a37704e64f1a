// NN potential (Coulomb descriptor + MLP): TensorflowPots/call_sample_model.py:4-9.
#pragma once
#include "pvd_step.cuh"
static int nn_launch_soa(cudaStream_t, const double *, const DevState *, int, long long, double *, int)
{
    return pvd_fail(PVD_E_STATE, "NN potential: not built yet");
}

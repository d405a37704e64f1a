// Importance sampling kernels (pyvibdmc.py:549-612, imp_samp.py:21-76).
#pragma once
#include "pvd_step.cuh"

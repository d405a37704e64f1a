// Continuous weighting (pyvibdmc.py:432-454 + _branch :340-356): types shared with the host code.
#pragma once
#include "pvd_step.cuh"

constexpr int PVD_HIST_BINS = 4096;
struct ContCand { double w; int idx; int pad; };
struct ContWork { int n_kill; int n_cand; double edge; int fallback; int pad; };

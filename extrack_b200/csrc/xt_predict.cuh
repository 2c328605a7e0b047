// K3 — state annotation kernel (predict_Bs).  Placeholder until the fit path is parity-green.
#pragma once
#include "xt_common.cuh"

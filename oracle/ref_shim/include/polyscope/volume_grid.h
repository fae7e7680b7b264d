// Shim header: see shm_ref_shim.h (oracle/ref_shim) -- NOT the real library.
#pragma once
#include "shm_ref_shim.h"

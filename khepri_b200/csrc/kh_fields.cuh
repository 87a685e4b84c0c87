// Field reconstruction kernels (crystal.py:234-343, fields.py, fourier.py:136-142) -- see kh_fields_impl below.
#pragma once
#include "kh_common.cuh"

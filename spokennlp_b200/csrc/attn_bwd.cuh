#pragma once
#include "ptx.cuh"

#pragma once
#include "uchar.h"

// b3d_internal.h — host-side declarations shared by the translation units of libb3d.so (not part of the public ABI).
#pragma once
#include <stdlib.h>
#include <string.h>

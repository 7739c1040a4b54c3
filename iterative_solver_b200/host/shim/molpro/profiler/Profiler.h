#include <molpro/Profiler.h>

#pragma once
// Empty on purpose: the reference headers include gsl-lite but the translation units compiled for the
// oracle (ray_tracing.cpp, bounding_volume_hierarchy.cpp, shadow.cpp) never use it.

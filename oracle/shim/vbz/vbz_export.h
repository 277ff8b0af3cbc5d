/* Stand-in for the header cmake's generate_export_header() would produce
 * (reference vbz/CMakeLists.txt:16). */
#pragma once
#define VBZ_EXPORT __attribute__((visibility("default")))
#define VBZ_NO_EXPORT __attribute__((visibility("hidden")))

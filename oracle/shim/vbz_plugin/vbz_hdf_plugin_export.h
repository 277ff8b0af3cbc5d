/* Stand-in for the generated export header (reference vbz_plugin/CMakeLists.txt:9). */
#pragma once
#define VBZ_HDF_PLUGIN_EXPORT __attribute__((visibility("default")))

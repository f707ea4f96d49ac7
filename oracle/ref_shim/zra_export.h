/* Stand-in for the export header the reference's CMake would generate
 * (GENERATE_EXPORT_HEADER, /root/reference/CMakeLists.txt:51). Written for the
 * oracle build only: every reference symbol gets default visibility so the
 * test harness can bind it through ctypes. */
#ifndef ZRA_EXPORT_H
#define ZRA_EXPORT_H
#define ZRA_EXPORT __attribute__((visibility("default")))
#define ZRA_NO_EXPORT __attribute__((visibility("hidden")))
#define ZRA_DEPRECATED __attribute__((__deprecated__))
#endif

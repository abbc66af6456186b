/* Minimal autoconf config.h stand-in for building reference TUs in oracle/_ref. */
#ifndef ORACLE_SHIM_CONFIG_H
#define ORACLE_SHIM_CONFIG_H
#define PACKAGE_VERSION "oracle-ref"
#define HAVE_GETRUSAGE 1
#define GIT_HEADHASH "unknown"
/* HAVE_LONG_DOUBLE_WIDER (configure.ac:211) is deliberately absent: mosaic_util.c and
 * create_gnomonic_cubic_grid.c test it but never include config.h, so a real autotools build compiles their
 * plain-double branches (only the acosl() call in spherical_angle, mosaic_util.c:834, is x87). */
#endif

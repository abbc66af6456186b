/* Minimal autoconf config.h stand-in for building reference TUs in oracle/_ref. */
#ifndef ORACLE_SHIM_CONFIG_H
#define ORACLE_SHIM_CONFIG_H
#define PACKAGE_VERSION "oracle-ref"
#define HAVE_GETRUSAGE 1
#define GIT_HEADHASH "unknown"
#endif

/* ORACLE / TEST INFRASTRUCTURE.  Version macros the reference stringizes at
 * algebra/_common/lin_sys/qdldl/qdldl_interface.c:389.  Restated algorithm of v0.1.8. */
#ifndef QDLDL_VERSION_H
#define QDLDL_VERSION_H
#define QDLDL_VERSION_MAJOR 0
#define QDLDL_VERSION_MINOR 1
#define QDLDL_VERSION_PATCH 8
#endif

/* Compatibility include: lets a C program written against the reference's
 * "structured/StructuredMatrix.h" (e.g. examples/dense/dstructured.c) compile
 * unchanged against the engine: add -I<repo>/include/compat and link
 * -lstrumpack_b200.  Everything is declared in sb200_structured.h. */
#ifndef SB200_COMPAT_STRUCTURED_MATRIX_H
#define SB200_COMPAT_STRUCTURED_MATRIX_H
#include "../../sb200_structured.h"
#endif

/* What CMake's FortranCInterface would emit for gfortran-style mangling
 * (reference CMakeLists.txt:61-65). TEST INFRASTRUCTURE ONLY. */
#ifndef STRUMPACK_FC_HEADER_INCLUDED
#define STRUMPACK_FC_HEADER_INCLUDED
#define STRUMPACK_FC_GLOBAL(name,NAME) name##_
#define STRUMPACK_FC_GLOBAL_(name,NAME) name##_
#endif

// Stand-in for <osqp/osqp.h> (OSQP is absent from the image); the reference only needs the header to exist.
#pragma once
typedef double c_float;
typedef long long c_int;
#define OSQP_INFTY ((c_float)1e30)

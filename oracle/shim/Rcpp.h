// see RcppEigen.h in this directory
#include "RcppEigen.h"

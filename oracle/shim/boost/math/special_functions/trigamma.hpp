#include "digamma.hpp"

// stand-in header, see xtl_shim_all.hpp
#include "xtl_shim_all.hpp"

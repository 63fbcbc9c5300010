// cvshim highgui (TEST INFRASTRUCTURE ONLY) -- nothing from highgui is used on the path.
#ifndef CVSHIM_HIGHGUI_HPP
#define CVSHIM_HIGHGUI_HPP
#include "opencv2/core/core.hpp"
#endif

// TEST INFRASTRUCTURE -- stand-in, see opencv.hpp
#pragma once
#include "opencv.hpp"

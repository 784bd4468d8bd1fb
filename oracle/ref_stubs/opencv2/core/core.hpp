// see opencv2/opencv.hpp in this directory
#pragma once
#include <opencv2/opencv.hpp>

// empty stand-in: corridor.h includes OpenCV but declares nothing with it
#pragma once

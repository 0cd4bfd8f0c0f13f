/* The reference writes #include <opencv2\opencv.hpp> (Windows path separator): on Linux that is a file whose NAME contains a backslash. */
#include "opencv2/opencv.hpp"

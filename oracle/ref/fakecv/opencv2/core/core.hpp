/* oracle/ref/fakecv -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.  This image has no OpenCV C++ headers: the reference's line_lbd sources include
 * <opencv2/...>, and this directory on the include path hands them oracle/ref/minicv.hpp instead (see that file). */
#include "../../../minicv.hpp"

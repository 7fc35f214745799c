/* oracle/ref/fakecv -- CPU ORACLE, TEST INFRASTRUCTURE ONLY: hands oracle/ref/minicv.hpp to reference sources that include OpenCV headers (see that file). */
#include "../../../minicv.hpp"

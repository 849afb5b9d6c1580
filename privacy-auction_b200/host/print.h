// Console macros with the names of the reference's include/print.h.
#ifndef PA_HOST_PRINT_H
#define PA_HOST_PRINT_H
#include <iostream>

namespace pa_host {
inline bool &quiet() {
  static bool q = false;
  return q;
}
}  // namespace pa_host

#define PA_RULE "\n================================================\n"
#define PRINT_MESSAGE(msg)                                                                      \
  do {                                                                                          \
    if (!pa_host::quiet()) std::cout << PA_RULE << __FILE__ << ":" << __LINE__ << ": \n" << msg << PA_RULE << std::endl; \
  } while (0)
#define PRINT_ERROR(msg) \
  std::cerr << PA_RULE << "\x1b[31m[ERROR] \x1b[0m" << __FILE__ << ":" << __LINE__ << ": \n" << msg << PA_RULE << std::endl
#define PRINT_INFO(msg) \
  std::cout << PA_RULE << "\x1b[34m[INFO] \x1b[0m" << __FILE__ << ":" << __LINE__ << ": \n" << msg << PA_RULE << std::endl

#endif

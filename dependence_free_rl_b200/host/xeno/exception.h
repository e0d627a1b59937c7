// xeno/exception.h -- error convention of the reference (xeno/exception.h:12-23): a runtime_error
// that remembers where it was thrown. Every failure of the C ABI (include/dfrl.h) surfaces as one.
#ifndef XENO_EXCEPTION_
#define XENO_EXCEPTION_

#include <source_location>
#include <stdexcept>
#include <string>

namespace xeno {

class error : public std::runtime_error {
public:
  explicit error(const std::string &what, std::source_location where = std::source_location::current())
      : std::runtime_error(std::string(where.file_name()) + ":" + std::to_string(where.line()) + ": " + what),
        where_(where) {}
  const std::source_location &where() const { return where_; }

private:
  std::source_location where_;
};

} // namespace xeno

#endif // XENO_EXCEPTION_

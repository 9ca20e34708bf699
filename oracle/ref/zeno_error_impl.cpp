// TEST INFRASTRUCTURE ONLY -- the out-of-line members of zeno/utils/Error.h (declared there, defined in libzeno, which
// oracle/_ref does not link). The reference's IObject::as<T>() reports a wrong socket type through them
// (zeno/utils/safe_dynamic_cast.h); the node harnesses only need them to exist and to carry a message.
#include <zeno/utils/Error.h>

#include <exception>

namespace zeno {
Error::Error(std::string_view m) noexcept : message(m) {}
Error::~Error() noexcept = default;
std::string const& Error::what() const noexcept { return message; }
StdError::StdError(std::exception_ptr&& e) noexcept : Error("std::exception"), eptr(std::move(e)) {}
StdError::~StdError() noexcept = default;
TypeError::TypeError(std::type_info const& e, std::type_info const& g, std::string_view h) noexcept
    : Error(std::string("expect ") + e.name() + ", got " + g.name() + " (" + std::string(h) + ")"), expect(e), got(g), hint(h) {}
TypeError::~TypeError() noexcept = default;
KeyError::KeyError(std::string_view k, std::string_view h) noexcept : Error(std::string("invalid key ") + std::string(k)), key(k), hint(h) {}
KeyError::~KeyError() noexcept = default;
IndexError::IndexError(size_t i, size_t m, std::string_view h) noexcept : Error("index out of range"), index(i), maxRange(m), hint(h) {}
IndexError::~IndexError() noexcept = default;
UnimplError::UnimplError(std::string_view h) noexcept : Error("not implemented"), hint(h) {}
UnimplError::~UnimplError() noexcept = default;
ErrorException::ErrorException(std::shared_ptr<Error>&& e) noexcept : err(std::move(e)) {}
ErrorException::~ErrorException() noexcept = default;
char const* ErrorException::what() const noexcept { return err ? err->what().c_str() : "zeno error"; }
std::shared_ptr<Error> ErrorException::getError() const noexcept { return err; }
}  // namespace zeno

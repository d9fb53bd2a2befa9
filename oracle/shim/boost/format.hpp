// Minimal stand-in for <boost/format.hpp> -- TEST INFRASTRUCTURE ONLY.
// The reference's examples/implicit/implicit.hpp builds an output file name with
// boost::format("out_%d.data") % num; the oracle driver never reaches that branch
// (save_every is huge) but the expression must compile.
#ifndef ADSB_SHIM_BOOST_FORMAT_HPP
#define ADSB_SHIM_BOOST_FORMAT_HPP

#include <sstream>
#include <string>

namespace boost {

class format {
    std::string fmt_;
    std::ostringstream args_;

public:
    explicit format(std::string f) : fmt_{std::move(f)} { }
    format(const format& o) : fmt_{o.fmt_} { args_ << o.args_.str(); }

    template <typename T>
    format& operator%(const T& v) {
        args_ << '_' << v;
        return *this;
    }

    std::string str() const { return fmt_ + args_.str(); }
};

inline std::string str(const format& f) { return f.str(); }

}  // namespace boost

#endif

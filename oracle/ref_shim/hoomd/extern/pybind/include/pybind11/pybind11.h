// Stand-in for pybind11 when the reference's sources are compiled for the oracle cross-check (oracle/Makefile, target _ref):
// the export_*() functions of the reference only have to compile, they are never called.  TEST INFRASTRUCTURE ONLY.
#pragma once
namespace pybind11 {
struct module {};
template <class T> struct base {};
template <class... A> struct init {};
template <class... T> struct class_ {
    template <class... A> class_(A&&...) {}
    template <class... A> class_& def(A&&...) { return *this; }
    template <class... A> class_& def_readwrite(A&&...) { return *this; }
    template <class... A> class_& def_property(A&&...) { return *this; }
};
template <class T> struct enum_ {
    template <class... A> enum_(A&&...) {}
    template <class... A> enum_& value(A&&...) { return *this; }
    enum_& export_values() { return *this; }
};
template <class T> void bind_vector(module&, const char*) {}
}  // namespace pybind11

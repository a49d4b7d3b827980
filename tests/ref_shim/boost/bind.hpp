// stand-in for the two uses of boost::bind in the reference:
//   list.sort(boost::bind(&Feature_Alignment::CellComparator, _1, _2))                                   ref: src/Feature_alignment.cpp:88
//   list.sort(boost::bind(&std::pair<KeyFrame*, double>::second, _1) < boost::bind(&...::second, _2))     ref: src/Tracking.cpp:267-268,340-341
// i.e. a binary function bound to (_1, _2), and a data member projected from _1 / _2 with boost's operator< between two binders.
#ifndef MINI_BOOST_BIND_H
#define MINI_BOOST_BIND_H
namespace boost {
template <int N> struct arg {};
namespace mini {
template <class F> struct bound2 {            // f(_1, _2)
    F f;
    template <class A, class B> auto operator()(A& a, B& b) const -> decltype(f(a, b)) { return f(a, b); }
};
template <class T, class M, int N> struct member_of {   // (_N).*pm
    M T::*pm;
    template <class A, class B> const M& operator()(const A& a, const B& b) const { return pick(a, b, arg<N>()).*pm; }
private:
    template <class A, class B> static const A& pick(const A& a, const B&, arg<1>) { return a; }
    template <class A, class B> static const B& pick(const A&, const B& b, arg<2>) { return b; }
};
template <class L, class R> struct less2 {
    L l; R r;
    template <class A, class B> bool operator()(const A& a, const B& b) const { return l(a, b) < r(a, b); }
};
template <class T, class M, int N1, class T2, class M2, int N2>
less2<member_of<T, M, N1>, member_of<T2, M2, N2> > operator<(const member_of<T, M, N1>& l, const member_of<T2, M2, N2>& r)
{
    less2<member_of<T, M, N1>, member_of<T2, M2, N2> > x = { l, r };
    return x;
}
}  // namespace mini
template <class R, class A, class B> mini::bound2<R (*)(A, B)> bind(R (*f)(A, B), arg<1>, arg<2>)
{
    mini::bound2<R (*)(A, B)> x = { f };
    return x;
}
template <class T, class M, int N> mini::member_of<T, M, N> bind(M T::*pm, arg<N>)
{
    mini::member_of<T, M, N> x = { pm };
    return x;
}
}  // namespace boost
static boost::arg<1> _1;
static boost::arg<2> _2;
#endif

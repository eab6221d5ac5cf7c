#pragma once
namespace boost {
template <class T> class intrusive_ptr {
  T* p;
 public:
  intrusive_ptr() : p(0) {}
  intrusive_ptr(T* q, bool add_ref = true) : p(q) { if (p && add_ref) intrusive_ptr_add_ref(p); }
  intrusive_ptr(const intrusive_ptr& o) : p(o.p) { if (p) intrusive_ptr_add_ref(p); }
  ~intrusive_ptr() { if (p) intrusive_ptr_release(p); }
  intrusive_ptr& operator=(const intrusive_ptr& o) { intrusive_ptr t(o); T* q = p; p = t.p; t.p = q; return *this; }
  T* get() const { return p; }
  T& operator*() const { return *p; }
  T* operator->() const { return p; }
  operator bool() const { return p != 0; }
};
template <class T, class U> bool operator==(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) { return a.get() == b.get(); }
template <class T, class U> bool operator!=(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) { return a.get() != b.get(); }
}

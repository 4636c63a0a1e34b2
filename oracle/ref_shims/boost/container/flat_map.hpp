// Minimal stand-in for boost::container::flat_map (Boost 1.84 is a Conan dependency of the
// reference and is not present under /root/reference).  TEST INFRASTRUCTURE ONLY: it exists so the
// reference's own sources can be compiled into oracle/_ref/ as the parity oracle.  Only the subset
// of the API that the reference's hot-path sources use is provided (sorted-vector semantics).
#ifndef DPHY_ORACLE_SHIM_BOOST_FLAT_MAP_HPP_
#define DPHY_ORACLE_SHIM_BOOST_FLAT_MAP_HPP_

#include <algorithm>
#include <functional>
#include <memory>
#include <utility>
#include <vector>

namespace boost {
namespace container {

struct ordered_unique_range_t {};

template<typename Key, typename T, typename Compare = std::less<Key>,
         typename Allocator = std::allocator<std::pair<Key, T>>>
class flat_map {
 public:
  using key_type = Key;
  using mapped_type = T;
  using value_type = std::pair<Key, T>;
  using allocator_type =
      typename std::allocator_traits<Allocator>::template rebind_alloc<value_type>;
  using sequence_type = std::vector<value_type, allocator_type>;
  using iterator = value_type*;               // raw pointers: iterators of maps with different
  using const_iterator = const value_type*;   // allocators must be the same type (interval_set.h:258-262)
  using size_type = typename sequence_type::size_type;
  using difference_type = typename sequence_type::difference_type;

  flat_map() = default;
  template<typename A>
  explicit flat_map(const A& alloc) : seq_(allocator_type(alloc)) {}

  iterator begin() { return seq_.data(); }
  iterator end() { return seq_.data() + seq_.size(); }
  const_iterator begin() const { return seq_.data(); }
  const_iterator end() const { return seq_.data() + seq_.size(); }
  const_iterator cbegin() const { return begin(); }
  const_iterator cend() const { return end(); }

  bool empty() const { return seq_.empty(); }
  size_type size() const { return seq_.size(); }
  void clear() { seq_.clear(); }
  void reserve(size_type n) { seq_.reserve(n); }

  iterator lower_bound(const Key& k) {
    return std::lower_bound(begin(), end(), k,
                            [](const value_type& v, const Key& key) { return Compare{}(v.first, key); });
  }
  const_iterator lower_bound(const Key& k) const {
    return std::lower_bound(begin(), end(), k,
                            [](const value_type& v, const Key& key) { return Compare{}(v.first, key); });
  }
  iterator upper_bound(const Key& k) {
    return std::upper_bound(begin(), end(), k,
                            [](const Key& key, const value_type& v) { return Compare{}(key, v.first); });
  }
  const_iterator upper_bound(const Key& k) const {
    return std::upper_bound(begin(), end(), k,
                            [](const Key& key, const value_type& v) { return Compare{}(key, v.first); });
  }
  iterator find(const Key& k) {
    auto it = lower_bound(k);
    return (it != end() && !Compare{}(k, it->first)) ? it : end();
  }
  const_iterator find(const Key& k) const {
    auto it = lower_bound(k);
    return (it != end() && !Compare{}(k, it->first)) ? it : end();
  }
  bool contains(const Key& k) const { return find(k) != end(); }
  size_type count(const Key& k) const { return contains(k) ? 1 : 0; }

  template<typename P>
  std::pair<iterator, bool> insert(const P& v) {
    auto it = lower_bound(v.first);
    if (it != end() && !Compare{}(v.first, it->first)) { return {it, false}; }
    auto off = it - begin();
    seq_.insert(seq_.begin() + off, value_type(v.first, v.second));
    return {begin() + off, true};
  }
  std::pair<iterator, bool> insert(const value_type& v) { return insert<value_type>(v); }
  template<typename It>
  void insert(It first, It last) {
    for (; first != last; ++first) { insert(*first); }
  }
  template<typename It>
  void insert(ordered_unique_range_t, It first, It last) {
    if (seq_.empty()) {
      for (; first != last; ++first) { seq_.emplace_back(first->first, first->second); }
    } else {
      insert(first, last);
    }
  }
  template<typename M>
  std::pair<iterator, bool> insert_or_assign(const Key& k, M&& m) {
    auto it = lower_bound(k);
    if (it != end() && !Compare{}(k, it->first)) {
      it->second = std::forward<M>(m);
      return {it, false};
    }
    auto off = it - begin();
    seq_.insert(seq_.begin() + off, value_type(k, std::forward<M>(m)));
    return {begin() + off, true};
  }
  template<typename... Args>
  std::pair<iterator, bool> try_emplace(const Key& k, Args&&... args) {
    auto it = lower_bound(k);
    if (it != end() && !Compare{}(k, it->first)) { return {it, false}; }
    auto off = it - begin();
    seq_.insert(seq_.begin() + off, value_type(k, T(std::forward<Args>(args)...)));
    return {begin() + off, true};
  }

  size_type erase(const Key& k) {
    auto it = find(k);
    if (it == end()) { return 0; }
    seq_.erase(seq_.begin() + (it - begin()));
    return 1;
  }
  iterator erase(const_iterator pos) {
    auto off = pos - begin();
    seq_.erase(seq_.begin() + off);
    return begin() + off;
  }
  iterator erase(const_iterator first, const_iterator last) {
    auto off = first - begin();
    seq_.erase(seq_.begin() + off, seq_.begin() + (last - begin()));
    return begin() + off;
  }

  T& at(const Key& k) { return find(k)->second; }
  const T& at(const Key& k) const { return find(k)->second; }
  T& operator[](const Key& k) { return try_emplace(k).first->second; }

  sequence_type extract_sequence() {
    sequence_type out(std::move(seq_));
    seq_.clear();
    return out;
  }
  void adopt_sequence(ordered_unique_range_t, sequence_type&& s) { seq_ = std::move(s); }
  void adopt_sequence(sequence_type&& s) {
    seq_ = std::move(s);
    std::sort(seq_.begin(), seq_.end(),
              [](const value_type& a, const value_type& b) { return Compare{}(a.first, b.first); });
  }

  friend bool operator==(const flat_map& a, const flat_map& b) {
    return a.seq_.size() == b.seq_.size() && std::equal(a.seq_.begin(), a.seq_.end(), b.seq_.begin());
  }
  friend bool operator!=(const flat_map& a, const flat_map& b) { return !(a == b); }

 private:
  sequence_type seq_;
};

}  // namespace container
}  // namespace boost

#endif  // DPHY_ORACLE_SHIM_BOOST_FLAT_MAP_HPP_

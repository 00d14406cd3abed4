// H5Lite — a dependency-free writer/reader for the subset of HDF5 that fv2d's IOManager
// uses (reference IOManager.h:99-398 through HighFive/H5Easy): libhdf5 is not available in
// this environment, so the run.h5 container is produced directly from the HDF5 File Format
// Specification.
//
// What is written is what libhdf5 writes with its default ("earliest") format bounds — the
// settings H5Easy::File uses — so the files open in h5py / HighFive / h5dump:
//   * version 0 superblock, 8-byte offsets and lengths, group K values 4 / 16;
//   * old-style groups: symbol-table message -> v1 B-tree ("TREE", node type 0) -> symbol
//     table nodes ("SNOD") -> local heap ("HEAP") holding the link names;
//   * version 1 object headers; messages: dataspace v1, datatype v1 (IEEE f64 LE, i32 LE,
//     variable-length UTF-8 string), fill value v2, contiguous data layout v3, attribute v1,
//     symbol table;
//   * std::string attributes are variable-length strings in a global heap collection
//     ("GCOL"), like HighFive's createAttribute(name, std::string).
// The reader accepts the same structures as libhdf5 produces them (multi-level B-trees,
// object-header continuation blocks, dataspace v2, attribute v2/v3, fixed-length strings), so
// a run.h5 written by the reference itself can be used as a restart file.
//
// Appending (File::ReadWrite, the mode the reference reopens run.h5 with for every snapshot,
// IOManager.h:202-204): raw data and the headers of new objects are appended at the end of
// the file; the structures of a group whose membership changed (local heap, symbol nodes,
// B-tree, object header) are rewritten at the end of the file and the superblock's root entry
// is updated last.  Dataset contents are never moved.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fv2d
{
namespace h5lite
{

constexpr uint64_t kUndef = ~0ULL;

struct Error : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------------------- byte buffers

class Bytes
{
public:
  std::vector<uint8_t> v;
  void u8(unsigned x) { v.push_back((uint8_t)x); }
  void u16(unsigned x)
  {
    u8(x & 0xff);
    u8((x >> 8) & 0xff);
  }
  void u32(uint32_t x)
  {
    for (int k = 0; k < 4; ++k)
      u8((x >> (8 * k)) & 0xff);
  }
  void u64(uint64_t x)
  {
    for (int k = 0; k < 8; ++k)
      u8((unsigned)((x >> (8 * k)) & 0xff));
  }
  void raw(const void *p, size_t n)
  {
    const uint8_t *b = static_cast<const uint8_t *>(p);
    v.insert(v.end(), b, b + n);
  }
  void str(const char *s) { raw(s, std::strlen(s)); }
  void zeros(size_t n) { v.insert(v.end(), n, 0); }
  void pad8() { zeros((8 - v.size() % 8) % 8); }
  void append(const Bytes &o) { v.insert(v.end(), o.v.begin(), o.v.end()); }
  size_t size() const { return v.size(); }
};

inline uint64_t rd(const uint8_t *p, int n)
{
  uint64_t x = 0;
  for (int k = n - 1; k >= 0; --k)
    x = (x << 8) | p[k];
  return x;
}
inline size_t pad8(size_t n) { return (n + 7) & ~size_t(7); }

// ---------------------------------------------------------------------------- object model

struct Attribute
{
  enum Kind { Int32, Float64, String } kind = Int32;
  int32_t i = 0;
  double d  = 0.0;
  std::string s;
  // where a String value lives once it is on disk (global heap collection, object index)
  uint64_t gheap_addr = kUndef;
  uint32_t gheap_idx  = 0;
};

class File;

// A group or a dataset.  Groups own their children; names are kept sorted because the
// on-disk B-tree is ordered by name (and HighFive::getObjectName(index) indexes by name).
class Object
{
public:
  bool is_group = true;
  std::vector<std::pair<std::string, Attribute>> attrs; // creation order
  std::map<std::string, std::unique_ptr<Object>> children;
  // dataset: 1-D or n-D fp64, contiguous
  std::vector<uint64_t> dims;
  uint64_t data_addr = kUndef, data_bytes = 0;
  // on-disk location (kUndef = not written yet); groups also cache their B-tree / heap
  uint64_t header_addr = kUndef, btree_addr = kUndef, heap_addr = kUndef;
  bool dirty = true; // needs (re)writing at flush

  File *file = nullptr;

  // ---- the subset of the HighFive interface IOManager uses
  Object &createGroup(const std::string &name);
  Object &createDataSet(const std::string &name, const std::vector<double> &values);
  void createAttribute(const std::string &name, int value);
  void createAttribute(const std::string &name, double value);
  void createAttribute(const std::string &name, const std::string &value);
  bool hasAttribute(const std::string &name) const { return findAttr(name) != nullptr; }
  bool exist(const std::string &name) const { return children.count(name) != 0; }
  const Attribute &getAttribute(const std::string &name) const
  {
    const Attribute *a = findAttr(name);
    if (!a)
      throw Error("attribute '" + name + "' does not exist");
    return *a;
  }
  void readAttribute(const std::string &name, double &out) const
  {
    const Attribute &a = getAttribute(name);
    out                = a.kind == Attribute::Float64 ? a.d : (double)a.i;
  }
  void readAttribute(const std::string &name, int &out) const
  {
    const Attribute &a = getAttribute(name);
    out                = a.kind == Attribute::Int32 ? a.i : (int)a.d;
  }
  size_t getNumberObjects() const { return children.size(); }
  std::string getObjectName(size_t index) const
  {
    if (index >= children.size())
      throw Error("object index out of range");
    auto it = children.begin();
    std::advance(it, (long)index);
    return it->first;
  }
  // "a/b/c" relative to this group
  const Object &get(const std::string &path) const
  {
    const Object *o = this;
    size_t p        = 0;
    while (p < path.size())
    {
      size_t q = path.find('/', p);
      if (q == std::string::npos)
        q = path.size();
      if (q > p)
      {
        auto it = o->children.find(path.substr(p, q - p));
        if (it == o->children.end())
          throw Error("object '" + path + "' does not exist");
        o = it->second.get();
      }
      p = q + 1;
    }
    return *o;
  }
  const Object &getGroup(const std::string &path) const
  {
    const Object &o = get(path);
    if (!o.is_group)
      throw Error("'" + path + "' is not a group");
    return o;
  }
  std::vector<uint64_t> getShape(const std::string &path) const
  {
    const Object &o = get(path);
    if (o.is_group)
      throw Error("'" + path + "' is not a dataset");
    return o.dims;
  }
  std::vector<double> load(const std::string &path) const;

private:
  const Attribute *findAttr(const std::string &name) const
  {
    for (auto &kv : attrs)
      if (kv.first == name)
        return &kv.second;
    return nullptr;
  }
  void addAttr(const std::string &name, Attribute a);
  void checkNew(const std::string &name) const
  {
    if (!is_group)
      throw Error("not a group");
    if (name.empty() || name.find('/') != std::string::npos)
      throw Error("invalid link name '" + name + "'");
    if (children.count(name))
      throw Error("object '" + name + "' already exists");
  }
};

// ---------------------------------------------------------------------------- the file

class File : public Object
{
public:
  enum Mode { ReadOnly, ReadWrite, Truncate };
  static constexpr unsigned kLeafK = 4, kInternalK = 16; // libhdf5 defaults (H5Pset_sym_k)

  File(const std::string &path, Mode mode) : path_(path), mode_(mode)
  {
    file     = this;
    is_group = true;
    if (mode == Truncate)
    {
      fp_ = std::fopen(path.c_str(), "wb+");
      if (!fp_)
        throw Error("Unable to create file " + path);
      Bytes z;
      z.zeros(kSuperblockBytes);
      writeAt(0, z);
      eof_ = kSuperblockBytes;
    }
    else
    {
      fp_ = std::fopen(path.c_str(), mode == ReadOnly ? "rb" : "rb+");
      if (!fp_)
        throw Error("Unable to open file " + path);
      try
      {
        readSuperblock();
        parseObject(*this, header_addr);
      }
      catch (...)
      {
        std::fclose(fp_);
        fp_ = nullptr;
        throw;
      }
    }
  }
  File(const File &)            = delete;
  File &operator=(const File &) = delete;
  ~File()
  {
    try
    {
      close();
    }
    catch (...)
    {
    }
  }

  void flush()
  {
    if (mode_ == ReadOnly || !fp_)
      return;
    if (anyDirty(*this))
    {
      writeObject(*this);
      writeSuperblock();
    }
    std::fflush(fp_);
  }
  void close()
  {
    if (!fp_)
      return;
    flush();
    std::fclose(fp_);
    fp_ = nullptr;
  }
  const std::string &getName() const { return path_; }

  // ---- used by Object
  uint64_t appendData(const void *p, size_t n)
  {
    requireWritable();
    const uint64_t addr = alignEof();
    writeRaw(base_ + addr, p, n);
    eof_ = addr + n;
    return addr;
  }
  void readData(uint64_t addr, void *p, size_t n) const { readRaw(base_ + addr, p, n); }
  void requireWritable() const
  {
    if (mode_ == ReadOnly)
      throw Error("file " + path_ + " is open read-only");
    if (!fp_)
      throw Error("file " + path_ + " is closed");
  }

private:
  static constexpr size_t kSuperblockBytes = 96;
  std::string path_;
  Mode mode_;
  FILE *fp_      = nullptr;
  uint64_t base_ = 0; // file offset of address 0 (user block)
  uint64_t eof_  = 0; // end-of-file ADDRESS (relative to base_)
  unsigned leaf_k_ = kLeafK, internal_k_ = kInternalK;

  // ---- raw IO
  void writeRaw(uint64_t off, const void *p, size_t n)
  {
    if (fseeko(fp_, (off_t)off, SEEK_SET) != 0 || (n && std::fwrite(p, 1, n, fp_) != n))
      throw Error("write to " + path_ + " failed");
  }
  void readRaw(uint64_t off, void *p, size_t n) const
  {
    if (fseeko(fp_, (off_t)off, SEEK_SET) != 0 || (n && std::fread(p, 1, n, fp_) != n))
      throw Error("read from " + path_ + " failed (truncated file?)");
  }
  void writeAt(uint64_t addr, const Bytes &b) { writeRaw(base_ + addr, b.v.data(), b.size()); }
  std::vector<uint8_t> readAt(uint64_t addr, size_t n) const
  {
    if (addr == kUndef)
      throw Error("undefined address in " + path_);
    std::vector<uint8_t> b(n);
    readRaw(base_ + addr, b.data(), n);
    return b;
  }
  uint64_t alignEof()
  {
    const uint64_t a = (eof_ + 7) & ~7ULL;
    if (a != eof_)
    {
      Bytes z;
      z.zeros((size_t)(a - eof_));
      writeAt(eof_, z);
      eof_ = a;
    }
    return a;
  }
  uint64_t append(const Bytes &b)
  {
    const uint64_t addr = alignEof();
    writeAt(addr, b);
    eof_ = addr + b.size();
    return addr;
  }
  static bool anyDirty(const Object &o)
  {
    if (o.dirty)
      return true;
    for (auto &kv : o.children)
      if (anyDirty(*kv.second))
        return true;
    return false;
  }

  // ======================================================================== encoding

  static Bytes encDatatype(Attribute::Kind k)
  {
    Bytes b;
    switch (k)
    {
    case Attribute::Int32: // class 0 fixed-point, little endian, signed, 32 bits
      b.u8(0x10);
      b.u8(0x08);
      b.u8(0);
      b.u8(0);
      b.u32(4);
      b.u16(0);
      b.u16(32);
      break;
    case Attribute::Float64: // class 1, IEEE binary64 little endian
      b.u8(0x11);
      b.u8(0x20);
      b.u8(0x3f);
      b.u8(0);
      b.u32(8);
      b.u16(0);  // bit offset
      b.u16(64); // precision
      b.u8(52);  // exponent location
      b.u8(11);  // exponent size
      b.u8(0);   // mantissa location
      b.u8(52);  // mantissa size
      b.u32(1023);
      break;
    case Attribute::String: // class 9 variable-length string, null-terminated, UTF-8
      b.u8(0x19);
      b.u8(0x01); // type = string, padding = null terminate
      b.u8(0x01); // character set UTF-8
      b.u8(0);
      b.u32(16);
      // base type: class 3 string of size 1, null-terminated, UTF-8
      b.u8(0x13);
      b.u8(0x10);
      b.u8(0);
      b.u8(0);
      b.u32(1);
      break;
    }
    return b;
  }
  static Bytes encDataspace(const std::vector<uint64_t> &dims)
  {
    Bytes b;
    b.u8(1);
    b.u8((unsigned)dims.size());
    b.u8(0);
    b.u8(0);
    b.u32(0);
    for (uint64_t d : dims)
      b.u64(d);
    return b;
  }
  static void putMessage(Bytes &out, unsigned type, unsigned flags, const Bytes &body)
  {
    const size_t n = pad8(body.size());
    out.u16(type);
    out.u16((unsigned)n);
    out.u8(flags);
    out.zeros(3);
    out.append(body);
    out.zeros(n - body.size());
  }

  // one 4096-byte global heap collection holding a single object (index 1)
  void storeVlenString(Attribute &a)
  {
    const size_t len = a.s.size();
    const size_t coll = std::max<size_t>(4096, pad8(16 + 16 + pad8(len) + 16));
    Bytes g;
    g.str("GCOL");
    g.u8(1);
    g.zeros(3);
    g.u64(coll);
    g.u16(1); // object index
    g.u16(0); // reference count
    g.u32(0);
    g.u64(len);
    g.raw(a.s.data(), len);
    g.pad8();
    // object 0: the free space that remains (its size field counts the 16-byte header too)
    const size_t remaining = coll - g.size();
    g.u16(0);
    g.u16(0);
    g.u32(0);
    g.u64(remaining);
    g.zeros(coll - g.size());
    a.gheap_addr = append(g);
    a.gheap_idx  = 1;
  }

  Bytes encAttribute(const std::string &name, Attribute &a)
  {
    if (a.kind == Attribute::String && a.gheap_addr == kUndef)
      storeVlenString(a);
    const Bytes dt = encDatatype(a.kind), ds = encDataspace({});
    Bytes b;
    b.u8(1);
    b.u8(0);
    b.u16((unsigned)name.size() + 1);
    b.u16((unsigned)dt.size());
    b.u16((unsigned)ds.size());
    b.raw(name.c_str(), name.size() + 1);
    b.pad8();
    b.append(dt);
    b.pad8();
    b.append(ds);
    b.pad8();
    switch (a.kind)
    {
    case Attribute::Int32: b.u32((uint32_t)a.i); break;
    case Attribute::Float64:
    {
      uint64_t u;
      std::memcpy(&u, &a.d, 8);
      b.u64(u);
      break;
    }
    case Attribute::String:
      b.u32((uint32_t)a.s.size());
      b.u64(a.gheap_addr);
      b.u32(a.gheap_idx);
      break;
    }
    return b;
  }

  uint64_t writeHeader(const Bytes &messages, unsigned nmsg)
  {
    Bytes h;
    h.u8(1);
    h.u8(0);
    h.u16(nmsg);
    h.u32(1); // object reference count
    h.u32((uint32_t)messages.size());
    h.u32(0); // pad to 8
    h.append(messages);
    return append(h);
  }

  void writeDataset(Object &o)
  {
    Bytes m;
    unsigned n = 0;
    putMessage(m, 0x0001, 0, encDataspace(o.dims));
    ++n;
    putMessage(m, 0x0003, 1, encDatatype(Attribute::Float64));
    ++n;
    {
      Bytes f; // fill value v2: late allocation, write fill "if set", default fill value
      f.u8(2);
      f.u8(2);
      f.u8(2);
      f.u8(1);
      f.u32(0);
      putMessage(m, 0x0005, 1, f);
      ++n;
    }
    {
      Bytes l; // data layout v3, contiguous
      l.u8(3);
      l.u8(1);
      l.u64(o.data_bytes ? o.data_addr : kUndef);
      l.u64(o.data_bytes);
      putMessage(m, 0x0008, 0, l);
      ++n;
    }
    for (auto &kv : o.attrs)
    {
      putMessage(m, 0x000C, 0, encAttribute(kv.first, kv.second));
      ++n;
    }
    o.header_addr = writeHeader(m, n);
    o.dirty       = false;
  }

  // Writes (or rewrites) a group: children first, then local heap, symbol nodes, B-tree and
  // the object header.  The addresses of all pieces are computed before anything is
  // serialised so that sibling pointers and keys can be filled in.
  void writeObject(Object &o)
  {
    if (!o.is_group)
    {
      if (o.dirty || o.header_addr == kUndef)
        writeDataset(o);
      return;
    }
    bool need = o.dirty || o.header_addr == kUndef;
    for (auto &kv : o.children)
    {
      const uint64_t before = kv.second->header_addr;
      if (anyDirty(*kv.second))
        writeObject(*kv.second);
      if (kv.second->header_addr != before)
        need = true; // a child moved: this group's symbol table must point to the new header
    }
    if (!need)
      return;

    // ---- local heap: "" at offset 0, then the names, then one free block
    std::vector<uint64_t> name_off;
    Bytes heap_data;
    heap_data.zeros(8);
    for (auto &kv : o.children)
    {
      name_off.push_back(heap_data.size());
      heap_data.raw(kv.first.c_str(), kv.first.size() + 1);
      heap_data.pad8();
    }
    const uint64_t free_off = heap_data.size();
    const uint64_t free_len = std::max<uint64_t>(32, pad8(heap_data.size()) / 2);
    heap_data.u64(1); // H5HL_FREE_NULL: last free block
    heap_data.u64(free_len);
    heap_data.zeros((size_t)free_len - 16);

    // ---- leaves (symbol table nodes) and B-tree levels
    const size_t per_leaf = 2 * leaf_k_, per_node = 2 * internal_k_;
    const size_t nchild   = o.children.size();
    const size_t nleaf    = (nchild + per_leaf - 1) / per_leaf;
    const size_t snod_bytes = 8 + per_leaf * 40;
    const size_t node_bytes = 24 + (per_node + 1) * 8 + per_node * 8;

    // level structure: level[0] = B-tree nodes pointing to SNODs, level[k+1] points to level[k]
    std::vector<size_t> count; // nodes per level
    {
      size_t n = std::max<size_t>(1, (nleaf + per_node - 1) / per_node);
      count.push_back(n);
      while (n > 1)
      {
        n = (n + per_node - 1) / per_node;
        count.push_back(n);
      }
    }
    // address plan (relative to the block start): heap header, heap data, SNODs, B-tree nodes
    uint64_t cursor        = (eof_ + 7) & ~7ULL;
    const uint64_t a_heap  = cursor;
    const uint64_t a_hdata = a_heap + 32;
    cursor                 = a_hdata + heap_data.size();
    const uint64_t a_snod  = cursor;
    cursor += nleaf * snod_bytes;
    std::vector<uint64_t> a_level(count.size());
    for (size_t l = 0; l < count.size(); ++l)
    {
      a_level[l] = cursor;
      cursor += count[l] * node_bytes;
    }

    Bytes blk;
    blk.str("HEAP");
    blk.u8(0);
    blk.zeros(3);
    blk.u64(heap_data.size());
    blk.u64(free_off);
    blk.u64(a_hdata);
    blk.append(heap_data);

    // entries in name order
    std::vector<Object *> kids;
    for (auto &kv : o.children)
      kids.push_back(kv.second.get());
    // max-name key (heap offset) of every leaf
    std::vector<uint64_t> leaf_maxkey(nleaf);
    for (size_t lf = 0; lf < nleaf; ++lf)
    {
      const size_t lo = lf * per_leaf, hi = std::min(nchild, lo + per_leaf);
      blk.str("SNOD");
      blk.u8(1);
      blk.u8(0);
      blk.u16((unsigned)(hi - lo));
      for (size_t e = lo; e < hi; ++e)
      {
        const Object &k = *kids[e];
        blk.u64(name_off[e]);
        blk.u64(k.header_addr);
        if (k.is_group)
        {
          blk.u32(1); // cached symbol-table information
          blk.u32(0);
          blk.u64(k.btree_addr);
          blk.u64(k.heap_addr);
        }
        else
        {
          blk.u32(0);
          blk.u32(0);
          blk.zeros(16);
        }
      }
      blk.zeros((per_leaf - (hi - lo)) * 40);
      leaf_maxkey[lf] = name_off[hi - 1];
    }
    // B-tree levels
    std::vector<uint64_t> child_addr(nleaf), child_max = leaf_maxkey;
    for (size_t lf = 0; lf < nleaf; ++lf)
      child_addr[lf] = a_snod + lf * snod_bytes;
    for (size_t l = 0; l < count.size(); ++l)
    {
      std::vector<uint64_t> next_addr, next_max;
      const size_t nc = child_addr.size();
      for (size_t nd = 0; nd < count[l]; ++nd)
      {
        const size_t lo = nd * per_node, hi = std::min(nc, lo + per_node);
        const size_t used = hi > lo ? hi - lo : 0;
        blk.str("TREE");
        blk.u8(0);
        blk.u8((unsigned)l);
        blk.u16((unsigned)used);
        blk.u64(nd > 0 ? a_level[l] + (nd - 1) * node_bytes : kUndef);
        blk.u64(nd + 1 < count[l] ? a_level[l] + (nd + 1) * node_bytes : kUndef);
        blk.u64(lo > 0 ? child_max[lo - 1] : 0); // key 0: "" or the largest name to the left
        for (size_t c = lo; c < hi; ++c)
        {
          blk.u64(child_addr[c]);
          blk.u64(child_max[c]);
        }
        blk.zeros((per_node - used) * 16);
        next_addr.push_back(a_level[l] + nd * node_bytes);
        next_max.push_back(used ? child_max[hi - 1] : 0);
      }
      child_addr.swap(next_addr);
      child_max.swap(next_max);
    }
    const uint64_t got = append(blk);
    if (got != a_heap || eof_ != cursor)
      throw Error("internal error: group layout plan does not match what was written");
    o.heap_addr  = a_heap;
    o.btree_addr = a_level.back();

    // ---- object header: symbol table message + attributes
    Bytes m;
    unsigned n = 0;
    {
      Bytes st;
      st.u64(o.btree_addr);
      st.u64(o.heap_addr);
      putMessage(m, 0x0011, 0, st);
      ++n;
    }
    for (auto &kv : o.attrs)
    {
      putMessage(m, 0x000C, 0, encAttribute(kv.first, kv.second));
      ++n;
    }
    o.header_addr = writeHeader(m, n);
    o.dirty       = false;
  }

  void writeSuperblock()
  {
    Bytes s;
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    s.raw(sig, 8);
    s.u8(0); // superblock version
    s.u8(0); // free-space storage version
    s.u8(0); // root group symbol table entry version
    s.u8(0);
    s.u8(0); // shared header message format version
    s.u8(8); // size of offsets
    s.u8(8); // size of lengths
    s.u8(0);
    s.u16(leaf_k_);
    s.u16(internal_k_);
    s.u32(0);      // file consistency flags
    s.u64(0);      // base address
    s.u64(kUndef); // free-space info
    s.u64(eof_);   // end-of-file address
    s.u64(kUndef); // driver information block
    // root group symbol table entry
    s.u64(0);
    s.u64(header_addr);
    s.u32(1);
    s.u32(0);
    s.u64(btree_addr);
    s.u64(heap_addr);
    if (s.size() != kSuperblockBytes)
      throw Error("internal error: superblock size");
    writeRaw(base_, s.v.data(), s.size());
  }

  // ======================================================================== decoding

  void readSuperblock()
  {
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    fseeko(fp_, 0, SEEK_END);
    const uint64_t fsize = (uint64_t)ftello(fp_);
    uint64_t off         = 0;
    uint8_t head[8];
    for (;;)
    {
      if (off + 8 > fsize)
        throw Error(path_ + " is not an HDF5 file");
      readRaw(off, head, 8);
      if (std::memcmp(head, sig, 8) == 0)
        break;
      off = off == 0 ? 512 : off * 2;
    }
    std::vector<uint8_t> b(kSuperblockBytes + 4);
    readRaw(off, b.data(), std::min<uint64_t>(b.size(), fsize - off));
    const unsigned ver = b[8];
    if (ver > 1)
      throw Error(path_ + ": superblock version " + std::to_string(ver) +
                  " (written with libver=latest) is not supported");
    if (b[13] != 8 || b[14] != 8)
      throw Error(path_ + ": only 8-byte offsets and lengths are supported");
    leaf_k_     = (unsigned)rd(&b[16], 2);
    internal_k_ = (unsigned)rd(&b[18], 2);
    size_t p    = 24 + (ver == 1 ? 4 : 0);
    const uint64_t base = rd(&b[p], 8);
    eof_                = rd(&b[p + 16], 8);
    p += 32;
    header_addr = rd(&b[p + 8], 8);
    // the base address is relative to the start of the file; a user block shifts both
    base_ = (base == 0 && off != 0) ? off : base;
    if (mode_ != ReadOnly && (off != 0 || base != 0))
      throw Error(path_ + ": files with a user block can only be opened read-only");
    if (base_ + eof_ > fsize)
      throw Error(path_ + " is truncated: end-of-file address " + std::to_string(eof_) + " beyond file size " +
                  std::to_string(fsize));
  }

  struct Msg
  {
    unsigned type, flags;
    std::vector<uint8_t> body;
  };

  std::vector<Msg> readMessages(uint64_t addr) const
  {
    auto h = readAt(addr, 16);
    if (h[0] != 1)
      throw Error(path_ + ": object header version " + std::to_string(h[0]) + " is not supported");
    const unsigned nmsg = (unsigned)rd(&h[2], 2);
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, rd(&h[8], 4)}};
    std::vector<Msg> out;
    for (size_t bi = 0; bi < blocks.size() && out.size() < nmsg; ++bi)
    {
      auto blk = readAt(blocks[bi].first, (size_t)blocks[bi].second);
      size_t q = 0;
      while (q + 8 <= blk.size() && out.size() < nmsg)
      {
        Msg m;
        m.type         = (unsigned)rd(&blk[q], 2);
        const size_t n = (size_t)rd(&blk[q + 2], 2);
        m.flags        = blk[q + 4];
        if (q + 8 + n > blk.size())
          throw Error(path_ + ": object header message overruns its block");
        m.body.assign(blk.begin() + (long)q + 8, blk.begin() + (long)(q + 8 + n));
        q += 8 + n;
        if (m.type == 0x0010 && m.body.size() >= 16)
          blocks.push_back({rd(&m.body[0], 8), rd(&m.body[8], 8)});
        out.push_back(std::move(m));
      }
    }
    return out;
  }

  struct TypeInfo
  {
    unsigned cls = 0, size = 0;
    bool is_signed = true, vlen_string = false;
    size_t encoded = 0;
  };
  TypeInfo decDatatype(const uint8_t *m, size_t n) const
  {
    if (n < 8)
      throw Error(path_ + ": short datatype message");
    TypeInfo t;
    t.cls  = m[0] & 0x0f;
    t.size = (unsigned)rd(m + 4, 4);
    if (m[1] & 1)
      if (t.cls == 0 || t.cls == 1)
        throw Error(path_ + ": big-endian data is not supported");
    switch (t.cls)
    {
    case 0:
      t.is_signed = (m[1] & 0x08) != 0;
      t.encoded   = 12;
      break;
    case 1:
      if (!((t.size == 8 && rd(m + 10, 2) == 64 && m[12] == 52 && m[13] == 11 && m[15] == 52) ||
            (t.size == 4 && rd(m + 10, 2) == 32 && m[12] == 23 && m[13] == 8 && m[15] == 23)))
        throw Error(path_ + ": non-IEEE floating-point type");
      t.encoded = 20;
      break;
    case 3: t.encoded = 8; break;
    case 9:
    {
      if ((m[1] & 0x0f) != 1)
        throw Error(path_ + ": variable-length sequences are not supported");
      t.vlen_string = true;
      t.encoded     = 8 + decDatatype(m + 8, n - 8).encoded;
      break;
    }
    default: throw Error(path_ + ": datatype class " + std::to_string(t.cls) + " is not supported");
    }
    return t;
  }
  // returns false for a null dataspace
  bool decDataspace(const uint8_t *m, size_t n, std::vector<uint64_t> &dims) const
  {
    dims.clear();
    if (n < 4)
      throw Error(path_ + ": short dataspace message");
    const unsigned ver = m[0], rank = m[1];
    size_t p;
    if (ver == 1)
      p = 8;
    else if (ver == 2)
    {
      if (m[3] == 2)
        return false;
      p = 4;
    }
    else
      throw Error(path_ + ": dataspace version " + std::to_string(ver));
    if (p + 8 * (size_t)rank > n)
      throw Error(path_ + ": short dataspace message");
    for (unsigned k = 0; k < rank; ++k)
      dims.push_back(rd(m + p + 8 * k, 8));
    return true;
  }
  std::string globalHeapString(uint64_t addr, uint32_t idx, uint32_t len) const
  {
    auto h = readAt(addr, 16);
    if (std::memcmp(h.data(), "GCOL", 4) != 0)
      throw Error(path_ + ": bad global heap signature");
    auto c   = readAt(addr, (size_t)rd(&h[8], 8));
    size_t q = 16;
    while (q + 16 <= c.size())
    {
      const unsigned oid = (unsigned)rd(&c[q], 2);
      const uint64_t osz = rd(&c[q + 8], 8);
      if (oid == idx)
        return std::string((const char *)&c[q + 16], std::min<uint64_t>(len, osz));
      if (oid == 0)
        break;
      q += 16 + pad8((size_t)osz);
    }
    throw Error(path_ + ": global heap object not found");
  }

  void decAttribute(const Msg &msg, Object &o) const
  {
    const auto &m      = msg.body;
    const unsigned ver = m.at(0);
    if (ver < 1 || ver > 3)
      throw Error(path_ + ": attribute message version " + std::to_string(ver));
    if (ver >= 2 && (m[1] & 3))
      throw Error(path_ + ": shared datatypes/dataspaces in attributes are not supported");
    const size_t nsz = (size_t)rd(&m[2], 2), tsz = (size_t)rd(&m[4], 2), ssz = (size_t)rd(&m[6], 2);
    size_t p      = 8 + (ver == 3 ? 1 : 0);
    auto step     = [&](size_t n) { return ver == 1 ? pad8(n) : n; };
    if (p + step(nsz) + step(tsz) + step(ssz) > m.size())
      throw Error(path_ + ": short attribute message");
    std::string name((const char *)&m[p], nsz ? nsz - 1 : 0);
    name = name.substr(0, name.find('\0'));
    p += step(nsz);
    const TypeInfo t = decDatatype(&m[p], tsz);
    p += step(tsz);
    std::vector<uint64_t> dims;
    const bool has = decDataspace(&m[p], ssz, dims);
    p += step(ssz);
    uint64_t count = 1;
    for (uint64_t d : dims)
      count *= d;
    if (!has || count != 1)
      return; // only scalar (or 1-element) attributes are mapped; others are ignored
    Attribute a;
    const uint8_t *v = &m[p];
    const size_t left = m.size() - p;
    if (t.vlen_string)
    {
      if (left < 16)
        throw Error(path_ + ": short attribute value");
      a.kind       = Attribute::String;
      a.gheap_addr = rd(v + 4, 8);
      a.gheap_idx  = (uint32_t)rd(v + 12, 4);
      a.s          = globalHeapString(a.gheap_addr, a.gheap_idx, (uint32_t)rd(v, 4));
    }
    else if (t.cls == 3)
    {
      a.kind = Attribute::String;
      a.s    = std::string((const char *)v, std::min<size_t>(left, t.size));
      a.s    = a.s.substr(0, a.s.find('\0'));
    }
    else if (t.cls == 1)
    {
      a.kind = Attribute::Float64;
      if (t.size == 8)
        std::memcpy(&a.d, v, 8);
      else
      {
        float f;
        std::memcpy(&f, v, 4);
        a.d = f;
      }
    }
    else
    {
      a.kind           = Attribute::Int32;
      const uint64_t u = rd(v, (int)std::min<unsigned>(t.size, 8));
      if (t.is_signed && t.size < 8 && (u >> (8 * t.size - 1)))
        a.i = (int32_t)(int64_t)(u | (~0ULL << (8 * t.size)));
      else
        a.i = (int32_t)u;
    }
    o.attrs.emplace_back(name, std::move(a));
  }

  std::string heapString(const std::vector<uint8_t> &heap_data, uint64_t off) const
  {
    if (off >= heap_data.size())
      throw Error(path_ + ": link name offset outside the local heap");
    const char *s = (const char *)&heap_data[(size_t)off];
    return std::string(s, strnlen(s, heap_data.size() - (size_t)off));
  }

  void walkGroupTree(uint64_t addr, const std::vector<uint8_t> &heap_data, Object &o, int depth)
  {
    if (depth > 64)
      throw Error(path_ + ": group B-tree too deep (cycle?)");
    auto h = readAt(addr, 8);
    if (std::memcmp(h.data(), "SNOD", 4) == 0)
    {
      const unsigned nsym = (unsigned)rd(&h[6], 2);
      auto e              = readAt(addr + 8, (size_t)nsym * 40);
      for (unsigned k = 0; k < nsym; ++k)
      {
        const std::string name = heapString(heap_data, rd(&e[40 * k], 8));
        auto child             = std::make_unique<Object>();
        child->file            = this;
        parseObject(*child, rd(&e[40 * k + 8], 8));
        o.children[name] = std::move(child);
      }
      return;
    }
    if (std::memcmp(h.data(), "TREE", 4) != 0 || h[4] != 0)
      throw Error(path_ + ": bad group B-tree node");
    const unsigned used = (unsigned)rd(&h[6], 2);
    auto body           = readAt(addr + 24, 8 + (size_t)used * 16);
    for (unsigned k = 0; k < used; ++k)
      walkGroupTree(rd(&body[8 + 16 * k], 8), heap_data, o, depth + 1);
  }

  void parseObject(Object &o, uint64_t addr)
  {
    o.header_addr = addr;
    o.dirty       = false;
    o.is_group    = false;
    bool has_layout = false;
    TypeInfo dtype;
    for (const Msg &m : readMessages(addr))
    {
      if (m.flags & 2)
      {
        if (m.type == 0x0001 || m.type == 0x0003 || m.type == 0x0008)
          throw Error(path_ + ": shared object header messages are not supported");
        continue;
      }
      switch (m.type)
      {
      case 0x0001: decDataspace(m.body.data(), m.body.size(), o.dims); break;
      case 0x0003: dtype = decDatatype(m.body.data(), m.body.size()); break;
      case 0x0008:
      {
        has_layout = true;
        if (m.body.size() >= 18 && m.body[0] == 3 && m.body[1] == 1)
        {
          o.data_addr  = rd(&m.body[2], 8);
          o.data_bytes = rd(&m.body[10], 8);
        }
        else
          o.data_addr = kUndef, o.data_bytes = kUndef; // layout this reader cannot load (chunked/compact)
        break;
      }
      case 0x000C: decAttribute(m, o); break;
      case 0x0011:
        o.is_group   = true;
        o.btree_addr = rd(&m.body[0], 8);
        o.heap_addr  = rd(&m.body[8], 8);
        break;
      default: break;
      }
    }
    if (o.is_group)
    {
      auto hh = readAt(o.heap_addr, 32);
      if (std::memcmp(hh.data(), "HEAP", 4) != 0)
        throw Error(path_ + ": bad local heap signature");
      auto heap_data = readAt(rd(&hh[24], 8), (size_t)rd(&hh[8], 8));
      walkGroupTree(o.btree_addr, heap_data, o, 0);
    }
    else
    {
      if (!has_layout)
        throw Error(path_ + ": object is neither a group nor a dataset");
      if (!(dtype.cls == 1 && dtype.size == 8))
        o.data_bytes = kUndef; // only fp64 datasets can be loaded
    }
  }

  friend class Object;
};

// ---------------------------------------------------------------------------- Object methods

inline void Object::addAttr(const std::string &name, Attribute a)
{
  file->requireWritable();
  if (findAttr(name))
    throw Error("attribute '" + name + "' already exists");
  attrs.emplace_back(name, std::move(a));
  dirty = true;
}
inline void Object::createAttribute(const std::string &name, int value)
{
  Attribute a;
  a.kind = Attribute::Int32;
  a.i    = value;
  addAttr(name, a);
}
inline void Object::createAttribute(const std::string &name, double value)
{
  Attribute a;
  a.kind = Attribute::Float64;
  a.d    = value;
  addAttr(name, a);
}
inline void Object::createAttribute(const std::string &name, const std::string &value)
{
  Attribute a;
  a.kind = Attribute::String;
  a.s    = value;
  addAttr(name, a);
}
inline Object &Object::createGroup(const std::string &name)
{
  file->requireWritable();
  checkNew(name);
  auto g      = std::make_unique<Object>();
  g->is_group = true;
  g->file     = file;
  Object &ref = *g;
  children[name] = std::move(g);
  dirty          = true;
  return ref;
}
inline Object &Object::createDataSet(const std::string &name, const std::vector<double> &values)
{
  file->requireWritable();
  checkNew(name);
  auto d        = std::make_unique<Object>();
  d->is_group   = false;
  d->file       = file;
  d->dims       = {(uint64_t)values.size()};
  d->data_bytes = values.size() * sizeof(double);
  if (!values.empty())
    d->data_addr = file->appendData(values.data(), values.size() * sizeof(double));
  Object &ref    = *d;
  children[name] = std::move(d);
  dirty          = true;
  return ref;
}
inline std::vector<double> Object::load(const std::string &path) const
{
  const Object &o = get(path);
  if (o.is_group)
    throw Error("'" + path + "' is not a dataset");
  if (o.data_bytes == kUndef)
    throw Error("dataset '" + path + "' is not a contiguous fp64 dataset");
  uint64_t n = 1;
  for (uint64_t d : o.dims)
    n *= d;
  if (n * sizeof(double) != o.data_bytes)
    throw Error("dataset '" + path + "': size mismatch");
  std::vector<double> out((size_t)n);
  if (n)
    file->readData(o.data_addr, out.data(), (size_t)o.data_bytes);
  return out;
}

} // namespace h5lite
} // namespace fv2d

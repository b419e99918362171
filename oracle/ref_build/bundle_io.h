// TEST INFRASTRUCTURE (oracle build only) -- not part of the product.
// "SMCB1" bundle: a flat list of named little-endian arrays, used to move inputs/outputs between
// pytest / bench.py (smcpp_b200/bundle.py reads and writes the same format) and the oracle binaries.
//
//   SMCB1\n
//   <name> <dtype> <ndim> <d0> ... <dN-1>\n      dtype in {i4,i8,f4,f8,u1}
//   <raw bytes>\n
//   ... repeated ...
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace smcb {

struct Array {
    std::string dtype;
    std::vector<long> shape;
    std::vector<char> data;
    long count() const { long c = 1; for (long d : shape) c *= d; return c; }
    template <typename T> const T *as() const { return reinterpret_cast<const T *>(data.data()); }
    template <typename T> T *as() { return reinterpret_cast<T *>(data.data()); }
};

inline int dtype_size(const std::string &dt)
{
    if (dt == "i4" || dt == "f4") return 4;
    if (dt == "i8" || dt == "f8") return 8;
    if (dt == "u1") return 1;
    throw std::runtime_error("bundle: unknown dtype " + dt);
}

struct Bundle {
    std::vector<std::string> order;
    std::map<std::string, Array> arrays;

    bool has(const std::string &name) const { return arrays.count(name) > 0; }
    const Array &get(const std::string &name) const
    {
        auto it = arrays.find(name);
        if (it == arrays.end()) throw std::runtime_error("bundle: missing entry " + name);
        return it->second;
    }
    double scalar_f8(const std::string &name) const { return get(name).as<double>()[0]; }
    int scalar_i4(const std::string &name) const { return get(name).as<int32_t>()[0]; }
    int scalar_i4(const std::string &name, int dflt) const { return has(name) ? scalar_i4(name) : dflt; }

    template <typename T>
    void put(const std::string &name, const std::string &dtype, const std::vector<long> &shape, const T *src)
    {
        Array a;
        a.dtype = dtype;
        a.shape = shape;
        if ((int)sizeof(T) != dtype_size(dtype)) throw std::runtime_error("bundle: dtype/size mismatch for " + name);
        a.data.resize(a.count() * sizeof(T));
        if (a.count()) std::memcpy(a.data.data(), src, a.data.size());
        if (!arrays.count(name)) order.push_back(name);
        arrays[name] = std::move(a);
    }
    void put_f8(const std::string &name, const std::vector<long> &shape, const double *src) { put<double>(name, "f8", shape, src); }
    void put_f4(const std::string &name, const std::vector<long> &shape, const float *src) { put<float>(name, "f4", shape, src); }
    void put_i4(const std::string &name, const std::vector<long> &shape, const int32_t *src) { put<int32_t>(name, "i4", shape, src); }
    void put_u1(const std::string &name, const std::vector<long> &shape, const uint8_t *src) { put<uint8_t>(name, "u1", shape, src); }

    void save(const std::string &path) const
    {
        FILE *f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("bundle: cannot write " + path);
        std::fputs("SMCB1\n", f);
        for (const std::string &name : order) {
            const Array &a = arrays.at(name);
            std::ostringstream hdr;
            hdr << name << " " << a.dtype << " " << a.shape.size();
            for (long d : a.shape) hdr << " " << d;
            hdr << "\n";
            std::fputs(hdr.str().c_str(), f);
            if (!a.data.empty()) std::fwrite(a.data.data(), 1, a.data.size(), f);
            std::fputc('\n', f);
        }
        std::fclose(f);
    }

    static Bundle load(const std::string &path)
    {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("bundle: cannot read " + path);
        Bundle b;
        char line[4096];
        if (!std::fgets(line, sizeof line, f) || std::strncmp(line, "SMCB1", 5) != 0) {
            std::fclose(f);
            throw std::runtime_error("bundle: bad magic in " + path);
        }
        while (std::fgets(line, sizeof line, f)) {
            std::istringstream hdr(line);
            std::string name;
            Array a;
            int ndim = 0;
            if (!(hdr >> name >> a.dtype >> ndim)) break;
            for (int i = 0; i < ndim; ++i) { long d; hdr >> d; a.shape.push_back(d); }
            size_t nbytes = (size_t)a.count() * dtype_size(a.dtype);
            a.data.resize(nbytes);
            if (nbytes && std::fread(a.data.data(), 1, nbytes, f) != nbytes) {
                std::fclose(f);
                throw std::runtime_error("bundle: truncated entry " + name);
            }
            std::fgetc(f);  // trailing newline
            b.order.push_back(name);
            b.arrays[name] = std::move(a);
        }
        std::fclose(f);
        return b;
    }
};

}  // namespace smcb

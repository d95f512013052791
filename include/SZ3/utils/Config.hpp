// SZ3/utils/Config.hpp -- drop-in replacement header (sz3_b200).
//
// Source-compatible with the reference's SZ3::Config (include/SZ3/utils/Config.hpp:54-78 enums, :146-177 constructor
// and setDims, :185-310 INI load/save, :312-413 save/load, :418-478 print/size_est/fields): same public field names,
// types and defaults, same methods, same serialized blob (the blob itself is produced by libsz3b200's
// sz3b_config_save / sz3b_config_load so there is one implementation of the byte layout).  It crosses the extern "C"
// boundary as the POD mirror sz3b_config (include/sz3b.h).
#ifndef SZ3_CONFIG_HPP
#define SZ3_CONFIG_HPP

#include <algorithm>
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "SZ3/def.hpp"
#include "SZ3/version.hpp"
#include "sz3b.h"

#define SZ_FLOAT 0
#define SZ_DOUBLE 1
#define SZ_UINT8 2
#define SZ_INT8 3
#define SZ_UINT16 4
#define SZ_INT16 5
#define SZ_UINT32 6
#define SZ_INT32 7
#define SZ_UINT64 8
#define SZ_INT64 9

namespace SZ3 {

enum EB { EB_ABS, EB_REL, EB_PSNR, EB_L2NORM, EB_ABS_AND_REL, EB_ABS_OR_REL };
enum ALGO { ALGO_LORENZO_REG, ALGO_INTERP_LORENZO, ALGO_INTERP, ALGO_NOPRED, ALGO_LOSSLESS, ALGO_BIOMD, ALGO_BIOMDXTC };
enum INTERP_ALGO { INTERP_ALGO_LINEAR, INTERP_ALGO_CUBIC };

const std::map<std::string, ALGO> ALGO_MAP = {
    {"ALGO_LORENZO_REG", ALGO_LORENZO_REG}, {"ALGO_INTERP_LORENZO", ALGO_INTERP_LORENZO}, {"ALGO_INTERP", ALGO_INTERP},
    {"ALGO_NOPRED", ALGO_NOPRED},           {"ALGO_LOSSLESS", ALGO_LOSSLESS},             {"ALGO_BIOMD", ALGO_BIOMD},
    {"ALGO_BIOMDXTC", ALGO_BIOMDXTC}};
const std::map<std::string, EB> EB_MAP = {{"ABS", EB_ABS},   {"REL", EB_REL},
                                          {"PSNR", EB_PSNR}, {"NORM", EB_L2NORM},
                                          {"ABS_AND_REL", EB_ABS_AND_REL}, {"ABS_OR_REL", EB_ABS_OR_REL}};
const std::map<std::string, INTERP_ALGO> INTERP_ALGO_MAP = {{"INTERP_ALGO_LINEAR", INTERP_ALGO_LINEAR},
                                                            {"INTERP_ALGO_CUBIC", INTERP_ALGO_CUBIC}};

inline std::string to_lower(const std::string &s) {
    std::string out = s;
    std::transform(out.begin(), out.end(), out.begin(), [](unsigned char ch) { return static_cast<char>(::tolower(ch)); });
    return out;
}

// case-insensitive name -> enum value; an unknown name is an invalid argument
template <class EnumType, class Dst>
void match_enum(const std::string &input, const std::map<std::string, EnumType> &table, Dst &dst) {
    const std::string want = to_lower(input);
    for (const auto &kv : table)
        if (to_lower(kv.first) == want) {
            dst = static_cast<Dst>(kv.second);
            return;
        }
    throw std::invalid_argument("Invalid enum value: " + input);
}
template <class EnumType>
std::string enum_to_string(EnumType value, const std::map<std::string, EnumType> &table) {
    for (const auto &kv : table)
        if (kv.second == value) return kv.first;
    throw std::invalid_argument("Invalid enum value");
}

template <class T>
const char *enum2Str(T e) {
    if (std::is_same<T, ALGO>::value) {
        static const char *n[] = {"ALGO_LORENZO_REG", "ALGO_INTERP_LORENZO", "ALGO_INTERP", "ALGO_NOPRED", "ALGO_LOSSLESS",
                                  "ALGO_BIOMD", "ALGO_BIOMDXTC"};
        return n[static_cast<int>(e)];
    }
    if (std::is_same<T, INTERP_ALGO>::value) return static_cast<int>(e) ? "INTERP_ALGO_CUBIC" : "INTERP_ALGO_LINEAR";
    static const char *n[] = {"EB_ABS", "EB_REL", "EB_PSNR", "EB_L2NORM", "EB_ABS_AND_REL", "EB_ABS_OR_REL"};
    return n[static_cast<int>(e)];
}

class Config {
   public:
    template <class... Dims>
    Config(Dims... args) {
        dims = std::vector<size_t>{static_cast<size_t>(args)...};
        setDims(dims.begin(), dims.end());
    }

    // drops dimensions of size 1, resets predDim and the per-rank default block size
    template <class Iter>
    size_t setDims(Iter begin, Iter end) {
        std::vector<size_t> in(begin, end);
        dims.clear();
        for (size_t d : in)
            if (d > 1) dims.push_back(d);
        if (dims.empty()) dims.push_back(1);
        N = static_cast<char>(dims.size());
        num = std::accumulate(dims.begin(), dims.end(), static_cast<size_t>(1), std::multiplies<size_t>());
        predDim = static_cast<uint8_t>(N);
        blockSize = N == 1 ? 128 : (N == 2 ? 16 : 6);
        return num;
    }

    void loadcfg(const std::string &ini_file_path) {
        std::ifstream f(ini_file_path);
        if (!f.is_open()) throw std::runtime_error("Failed to open config file: " + ini_file_path);
        std::ostringstream buf;
        buf << f.rdbuf();
        load_ini(buf.str());
    }

    void load_ini(const std::string &ini_content) {
        std::istringstream ss(ini_content);
        std::string line, section;
        auto trim = [](std::string &s) {
            const char *ws = " \t\r\n";
            const size_t a = s.find_first_not_of(ws);
            if (a == std::string::npos) {
                s.clear();
                return;
            }
            s = s.substr(a, s.find_last_not_of(ws) - a + 1);
        };
        auto truth = [](const std::string &v) {
            const std::string l = to_lower(v);
            return l == "true" || l == "1" || l == "yes" || l == "on";
        };
        while (std::getline(ss, line)) {
            trim(line);
            if (line.empty() || line[0] == '#') continue;
            if (line[0] == '[') {
                section = to_lower(line.substr(1, line.find(']') - 1));
                continue;
            }
            const size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string key = line.substr(0, eq), val = line.substr(eq + 1);
            trim(key);
            trim(val);
            key = to_lower(key);
            if (section == "globalsettings") {
                if (key == "cmpralgo") match_enum(val, ALGO_MAP, cmprAlgo);
                else if (key == "errorboundmode") match_enum(val, EB_MAP, errorBoundMode);
                else if (key == "abserrorbound") absErrorBound = std::stod(val);
                else if (key == "relerrorbound") relErrorBound = std::stod(val);
                else if (key == "psnrerrorbound") psnrErrorBound = std::stod(val);
                else if (key == "l2normerrorbound") l2normErrorBound = std::stod(val);
                else if (key == "openmp") openmp = truth(val);
            } else if (section == "algosettings") {
                if (key == "lorenzo") lorenzo = truth(val);
                else if (key == "lorenzo2ndorder") lorenzo2 = truth(val);
                else if (key == "regression") regression = truth(val);
                else if (key == "regression2ndorder") regression2 = truth(val);
                else if (key == "interpolationalgo") match_enum(val, INTERP_ALGO_MAP, interpAlgo);
                else if (key == "interpolationdirection") interpDirection = static_cast<uint8_t>(std::stoi(val));
                else if (key == "blocksize") blockSize = std::stoi(val);
                else if (key == "quantizationbintotal") quantbinCnt = std::stoi(val);
                else if (key == "interpolationanchorstride") interpAnchorStride = std::stoi(val);
                else if (key == "interpolationalpha") interpAlpha = std::stod(val);
                else if (key == "interpolationbeta") interpBeta = std::stod(val);
            }
        }
    }

    std::string save_ini() const {
        std::ostringstream ss;
        auto tf = [](bool b) { return b ? "true" : "false"; };
        ss << "[GlobalSettings]\n"
           << "CmprAlgo = " << enum_to_string(static_cast<ALGO>(cmprAlgo), ALGO_MAP) << "\n"
           << "ErrorBoundMode = " << enum_to_string(static_cast<EB>(errorBoundMode), EB_MAP) << "\n"
           << "AbsErrorBound = " << absErrorBound << "\n"
           << "RelErrorBound = " << relErrorBound << "\n"
           << "PSNRErrorBound = " << psnrErrorBound << "\n"
           << "L2NormErrorBound = " << l2normErrorBound << "\n"
           << "OpenMP = " << tf(openmp) << "\n"
           << "\n[AlgoSettings]\n"
           << "Lorenzo = " << tf(lorenzo) << "\n"
           << "Lorenzo2ndOrder = " << tf(lorenzo2) << "\n"
           << "Regression = " << tf(regression) << "\n"
           << "Regression2ndOrder = " << tf(regression2) << "\n"
           << "BlockSize = " << blockSize << "\n"
           << "QuantizationBinTotal = " << quantbinCnt << "\n"
           << "InterpolationAlgo = " << enum_to_string(static_cast<INTERP_ALGO>(interpAlgo), INTERP_ALGO_MAP) << "\n"
           << "InterpolationDirection = " << static_cast<int>(interpDirection) << "\n"
           << "InterpolationAnchorStride = " << interpAnchorStride << "\n"
           << "InterpolationAlpha = " << interpAlpha << "\n"
           << "InterpolationBeta = " << interpBeta << "\n";
        return ss.str();
    }

    // POD mirror for the C ABI; `openmp` carries the slab count (0 = single stream)
    sz3b_config to_pod(int omp_slabs = 0) const {
        sz3b_config p;
        p.N = N;
        for (int i = 0; i < 4; i++) p.dims[i] = i < N ? dims[i] : 0;
        p.cmprAlgo = cmprAlgo;
        p.errorBoundMode = errorBoundMode;
        p.absErrorBound = absErrorBound;
        p.relErrorBound = relErrorBound;
        p.psnrErrorBound = psnrErrorBound;
        p.l2normErrorBound = l2normErrorBound;
        p.openmp = openmp ? (omp_slabs > 0 ? omp_slabs : 1) : 0;
        p.quantbinCnt = quantbinCnt;
        p.blockSize = blockSize;
        p.lorenzo = lorenzo;
        p.lorenzo2 = lorenzo2;
        p.regression = regression;
        p.regression2 = regression2;
        p.interpAlgo = interpAlgo;
        p.interpDirection = interpDirection;
        p.interpAnchorStride = interpAnchorStride;
        p.interpAlpha = interpAlpha;
        p.interpBeta = interpBeta;
        p.dataType = dataType;
        p.predDim = predDim;
        return p;
    }
    void from_pod(const sz3b_config &p) {
        N = static_cast<char>(p.N);
        dims.assign(p.dims, p.dims + p.N);
        num = std::accumulate(dims.begin(), dims.end(), static_cast<size_t>(1), std::multiplies<size_t>());
        cmprAlgo = static_cast<uint8_t>(p.cmprAlgo);
        errorBoundMode = static_cast<uint8_t>(p.errorBoundMode);
        absErrorBound = p.absErrorBound;
        relErrorBound = p.relErrorBound;
        psnrErrorBound = p.psnrErrorBound;
        l2normErrorBound = p.l2normErrorBound;
        openmp = p.openmp != 0;
        quantbinCnt = p.quantbinCnt;
        blockSize = p.blockSize;
        lorenzo = p.lorenzo != 0;
        lorenzo2 = p.lorenzo2 != 0;
        regression = p.regression != 0;
        regression2 = p.regression2 != 0;
        dataType = static_cast<uint8_t>(p.dataType);
        predDim = static_cast<uint8_t>(p.predDim);
    }

    // serialized blob appended to every stream; advances c
    size_t save(unsigned char *&c) const {
        const sz3b_config p = to_pod();
        const size_t n = sz3b_config_save(&p, c);
        c += n;
        return n;
    }
    void load(const unsigned char *&c) {
        sz3b_config p = to_pod();
        const size_t len = static_cast<size_t>(c[0]);
        if (sz3b_config_load(&p, c, len) != 0) throw std::invalid_argument(sz3b_last_error());
        from_pod(p);
        c += len;
    }

    void print() {
        std::cout << "===================== Begin SZ3 Configuration =====================\n";
        std::cout << "sz3MagicNumber = " << sz3MagicNumber << "\n";
        std::cout << "sz3DataVer = " << versionStr(sz3DataVer) << "\n";
        std::cout << "Dimensions =";
        for (size_t d : dims) std::cout << " " << d;
        std::cout << "\n" << save_ini();
        std::cout << "===================== End SZ3 Configuration =====================\n";
    }

    size_t size_est() const {
        std::vector<uchar> buf(sizeof(Config) + 1024);
        unsigned char *p = buf.data();
        return save(p);
    }

    uint32_t sz3MagicNumber = SZ3_MAGIC_NUMBER;
    uint32_t sz3DataVer = versionInt(SZ3_DATA_VER);
    char N = 0;
    std::vector<size_t> dims;
    size_t num = 0;
    uint8_t cmprAlgo = ALGO_INTERP_LORENZO;
    uint8_t errorBoundMode = EB_ABS;
    double absErrorBound = 1e-3;
    double relErrorBound = 0.0;
    double psnrErrorBound = 0.0;
    double l2normErrorBound = 0.0;
    bool openmp = false;
    int quantbinCnt = 65536;
    int blockSize = 0;
    uint8_t predDim = 0;
    uint8_t dataType = SZ_FLOAT;
    bool lorenzo = true;
    bool lorenzo2 = false;
    bool regression = true;
    bool regression2 = false;
    uint8_t interpAlgo = INTERP_ALGO_CUBIC;
    uint8_t interpDirection = 0;
    int interpAnchorStride = -1;
    double interpAlpha = 1.25;
    double interpBeta = 2.0;
};

}  // namespace SZ3
#endif

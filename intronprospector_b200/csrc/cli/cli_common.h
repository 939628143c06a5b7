/* cli_common.h -- shared pieces of the two drop-in command line tools. */
#ifndef IPX_CLI_COMMON_H
#define IPX_CLI_COMMON_H
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <iostream>
#include <stdexcept>
#include <string>
#include "ipx.h"

static const char *IPX_PACKAGE_NAME = "intronProspector";
static const char *IPX_PACKAGE_VERSION = "1.4.0";
static const char *IPX_PACKAGE_URL = "https://github.com/diekhans/intronProspector";

/* cmd_error (intronProspector.cc:63-67) */
[[noreturn]] static inline void cmd_error(const std::string &msg) {
    std::cerr << "Error: " << msg << std::endl;
    std::cerr << "Use --help to get usage" << std::endl;
    exit(1);
}

/* AutoGzipOutput (AutoGzip.hh:41-65): gzip when the name ends in ".gz" */
class OutFile {
    FILE *fp_ = nullptr;
    gzFile gz_ = nullptr;
    std::string name_;

    public:
    explicit OutFile(const std::string &name) : name_(name) {
        const bool is_gz = name.size() >= 3 && name.compare(name.size() - 3, 3, ".gz") == 0;
        if (is_gz) {
            gz_ = gzopen(name.c_str(), "wb");
            if (!gz_) throw std::runtime_error("Failed to open gzip file: " + name);
        } else {
            fp_ = fopen(name.c_str(), "wb");
            if (!fp_) throw std::runtime_error("Failed to open file: " + name);
        }
    }
    ~OutFile() {
        if (fp_) fclose(fp_);
        if (gz_) gzclose(gz_);
    }
    void write(const char *data, size_t len) {
        if (len == 0) return;
        if (fp_) {
            if (fwrite(data, 1, len, fp_) != len) throw std::runtime_error("write error: " + name_);
            fflush(fp_); /* the reference ends every line with std::endl */
        } else {
            size_t off = 0;
            while (off < len) {
                unsigned chunk = (unsigned)std::min<size_t>(len - off, 1u << 30);
                if (gzwrite(gz_, data + off, chunk) <= 0) throw std::runtime_error("write error: " + name_);
                off += chunk;
            }
        }
    }
    void write(const ipx_text &t) { write(t.data, t.len); }
};

/* toUnsigned / toFloat (strconvert.hh:14-47) */
static inline unsigned cli_to_unsigned(const char *str) {
    char *end;
    errno = 0;
    unsigned long lnum = strtoul(str, &end, 10);
    if (end == str || *end != '\0') throw std::runtime_error("Invalid unsigned integer \"" + std::string(str) + "\"");
    unsigned num = (unsigned)lnum;
    if (errno != 0 || (unsigned long)num != lnum) throw std::runtime_error("Unsigned integer out of range \"" + std::string(str) + "\"");
    return num;
}

static inline float cli_to_float(const char *str) {
    char *end;
    errno = 0;
    float num = strtof(str, &end);
    if (end == str || *end != '\0' || errno != 0) {
        if (errno != 0) throw std::runtime_error("Float out of range \"" + std::string(str) + "\"");
        throw std::runtime_error("Invalid float \"" + std::string(str) + "\"");
    }
    return num;
}

static inline void check_ipx(ipx_handle *h, int rc) {
    if (rc != IPX_OK) throw std::runtime_error(h ? ipx_last_error(h) : "junction engine failure");
}

#endif

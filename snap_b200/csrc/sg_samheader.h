// sg_samheader.h -- the header of a SAM / BAM file (host code; no kernel involved): what SAMFormat::writeHeader (reference SNAPLib/SAM.cpp:1203-1305)
// and BAMFormat::writeHeader (Bam.cpp:969-1030) put in front of the records when the input was FASTQ (ReaderContext::header == NULL):
//   @HD VN:1.6 GO:query | SO:coordinate, the read-group line, @PG ID:SNAP PN:SNAP CL:<command line> VN:<version>, one @SQ per contig in ORIGINAL
//   contig order (SN, LN = the contig's length without the chromosome padding, AH:* for ALT contigs);
//   BAM: "BAM\1", l_text, that text, n_ref, and per contig l_name (with the NUL), the name, l_ref.
// With it the device output stage makes whole files: header || records -> snapgpu_bgzf_[deflate_]device -> + the 28-byte end-of-file member.
#ifndef SG_SAMHEADER_H
#define SG_SAMHEADER_H

#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

struct SgHeaderContig { std::string name; int64_t length; bool isAlt; };      // by ORIGINAL contig number; length without padding

// The contigs in ORIGINAL order from an index's host-side tables: `start` = beginningLocation (ascending, internal order), `original[c]` = the contig's
// place in the FASTA (Genome::Contig::originalContigNumber), a contig's length = the distance to the next beginningLocation (the last: to
// countOfBases) minus the chromosome padding (Genome::Contig::length, SAM.cpp:1292).  Returns false if `original` is not a permutation.
inline bool sg_header_contigs(const std::vector<std::string> &names, const std::vector<int64_t> &start, const std::vector<uint8_t> &isAlt,
                              const std::vector<int32_t> &original, int64_t countOfBases, uint32_t chromosomePadding, std::vector<SgHeaderContig> *out)
{
    const size_t nc = start.size();
    if (names.size() != nc || original.size() != nc) return false;
    out->assign(nc, SgHeaderContig());
    std::vector<uint8_t> seen(nc, 0);
    for (size_t c = 0; c < nc; c++) {
        const int32_t o = original[c];
        if (o < 0 || (size_t)o >= nc || seen[(size_t)o]) return false;
        seen[(size_t)o] = 1;
        SgHeaderContig &h = (*out)[(size_t)o];
        h.name = names[c];
        h.length = (c + 1 < nc ? start[c + 1] : countOfBases) - start[c] - (int64_t)chromosomePadding;
        h.isAlt = c < isAlt.size() && isAlt[c] != 0;
    }
    return true;
}

// SAMFormat::writeHeader for FASTQ input.  rgLine NULL = the reference's fallback "@RG\tID:FASTQ\tSM:sample" (SAM.cpp:1234); stock `snap-aligner`
// passes AlignerOptions::rgLineContents, "@RG\tID:FASTQ\tPL:Illumina\tPU:pu\tLB:lb\tSM:sm" unless -rg / -R changed it.
inline std::string sg_sam_header_text(const std::vector<SgHeaderContig> &contigs, bool sorted, const char *commandLine, const char *version, const char *rgLine,
                                      bool omitSQLines = false)
{
    std::string h = "@HD\tVN:1.6\t";
    h += sorted ? "SO:coordinate" : "GO:query";
    h += "\n";
    h += rgLine ? rgLine : "@RG\tID:FASTQ\tSM:sample";
    h += "\n@PG\tID:SNAP\tPN:SNAP\tCL:";
    h += commandLine ? commandLine : "";
    h += "\tVN:";
    h += version ? version : "";
    h += "\n";
    if (!omitSQLines) {
        for (size_t i = 0; i < contigs.size(); i++) {
            char num[32];
            snprintf(num, sizeof(num), "%llu", (unsigned long long)contigs[i].length);
            h += "@SQ\tSN:"; h += contigs[i].name; h += "\tLN:"; h += num;
            if (contigs[i].isAlt) h += "\tAH:*";
            h += "\n";
        }
    }
    return h;
}

// BAMFormat::writeHeader: the uncompressed header block of a .bam
inline std::vector<uint8_t> sg_bam_header(const std::vector<SgHeaderContig> &contigs, bool sorted, const char *commandLine, const char *version, const char *rgLine)
{
    const std::string text = sg_sam_header_text(contigs, sorted, commandLine, version, rgLine, false);
    std::vector<uint8_t> o;
    auto put32 = [&o](int32_t v) { for (int k = 0; k < 4; k++) o.push_back((uint8_t)((uint32_t)v >> (8 * k))); };
    o.push_back('B'); o.push_back('A'); o.push_back('M'); o.push_back(1);
    put32((int32_t)text.size());
    o.insert(o.end(), text.begin(), text.end());
    put32((int32_t)contigs.size());
    for (size_t i = 0; i < contigs.size(); i++) {
        put32((int32_t)contigs[i].name.size() + 1);
        o.insert(o.end(), contigs[i].name.begin(), contigs[i].name.end());
        o.push_back(0);
        put32((int32_t)contigs[i].length);
    }
    return o;
}

#endif

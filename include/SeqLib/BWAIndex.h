// SeqLib::BWAIndex -- drop-in for SeqLib/BWAIndex.h:27-76, backed by the B200 engine:
// the FM-index lives in HBM (one contiguous image), the suffix sort behind
// ConstructIndex runs on the GPU.  Same methods, same exceptions.
#pragma once
#include <memory>
#include <ostream>
#include <string>
#include "SeqLib/UnalignedSequence.h"
#include "SeqLib/BamHeader.h"

struct b200_index;

namespace SeqLib {

class BWAIndex {
public:
    BWAIndex() = default;
    ~BWAIndex();
    BWAIndex(const BWAIndex &) = delete;
    BWAIndex &operator=(const BWAIndex &) = delete;

    bool IsEmpty() const noexcept { return idx_ == nullptr; }
    BamHeader HeaderFromIndex() const;
    int NumSequences() const;
    std::string ChrIDToName(int id) const;            ///< std::out_of_range on a bad id, std::runtime_error without an index
    std::string printSamHeader() const;
    void ConstructIndex(const UnalignedSequenceVector &refs);   ///< std::invalid_argument on an empty name or sequence
    void LoadIndex(const std::string &prefix);        ///< std::runtime_error when the files cannot be loaded
    void WriteIndex(const std::string &prefix) const; ///< .bwt .sa .pac .ann .amb in bwa's format
    friend std::ostream &operator<<(std::ostream &os, const BWAIndex &idx);

    b200_index *handle() const { return idx_; }       ///< C ABI handle (include/seqlib_b200.h)

private:
    b200_index *idx_ = nullptr;
    friend class BWAAligner;
};

using BWAIndexPtr = std::shared_ptr<BWAIndex>;

} // namespace SeqLib

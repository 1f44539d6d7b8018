// dropin_main.cpp — TEST INFRASTRUCTURE.  The replay loop of the reference's driver (eqf_vio/src/main.cpp:108-170:
// merge IMU and measurement rows by stamp, skip rows at or before startTime, after every measurement row write
// `time, stateEstimate()` and `time, filter` with the reference's stream formatting) instantiated for two filter
// classes that share one call surface:
//     --impl ref   the reference's own `class VIOFilter` (its unmodified sources, compiled in place by the Makefile)
//     --impl b200  `class VIOFilterB200` (include/eqf_vio_b200/VIOFilterB200.h) over libeqvio_b200.so
// Same input files, same settings, same output code path (the reference's operator<< for VIOState; each class's own
// operator<< for the internal row): tests/test_dropin.py compares the two outputs.
//
//     dropin_replay --impl ref|b200 IMU.csv MEAS.csv CONFIG.yaml OUT_STATE.csv OUT_FILTER.csv [--precision P] [--aux]
//
// The reference reads its CSV rows through CSVReader.h and its YAML through yaml-cpp; neither is usable here (the
// Eigen stand-in has no MatrixBase, yaml-cpp is absent), so rows are split on ',' and the config is read by a
// reader for the flat two-section subset EQVIO_config_template.yaml uses.  --precision replaces the reference's
// setprecision(5) (main.cpp:136,139) for parity checks beyond five digits; --aux also exercises the
// AuxiliaryFilterData constructor and initialiseFromIMUData / setAuxiliaryData.
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "eqf_vio/IMUVelocity.h"
#include "eqf_vio/VIOFilter.h"
#include "eqf_vio/VIOFilterSettings.h"
#include "eqf_vio/VisionMeasurement.h"
#include "eqf_vio_b200/VIOFilterB200.h"

namespace {

typedef std::vector<std::string> Row;

std::vector<Row> readRows(const std::string& path) {
    std::ifstream in(path);
    if (!in) { std::cerr << "cannot open " << path << std::endl; std::exit(2); }
    std::vector<Row> rows;
    std::string line;
    bool header = true;
    while (std::getline(in, line)) {
        if (header) { header = false; continue; }   // main.cpp:58,63 skip the header line
        if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
        Row r;
        std::stringstream ss(line);
        std::string cell;
        while (std::getline(ss, cell, ',')) r.push_back(cell);
        rows.push_back(r);
    }
    return rows;
}

IMUVelocity imuOf(const Row& row) {   // row layout of main.cpp:184-190
    IMUVelocity v;
    v.stamp = std::stod(row[0]);
    v.omega = Eigen::Vector3d(std::stod(row[1]), std::stod(row[2]), std::stod(row[3]));
    v.accel = Eigen::Vector3d(std::stod(row[4]), std::stod(row[5]), std::stod(row[6]));
    return v;
}

VisionMeasurement measOf(const Row& row) {   // row layout of main.cpp:192-203
    VisionMeasurement m;
    m.stamp = std::stod(row[0]);
    m.numberOfBearings = std::stoi(row[1]);
    m.bearings.resize(m.numberOfBearings);
    for (int i = 0; i < m.numberOfBearings; ++i) {
        const int j = 2 + 4 * i;
        m.bearings[i].id = std::stoi(row[j]);
        m.bearings[i].p = Eigen::Vector3d(std::stod(row[j + 1]), std::stod(row[j + 2]), std::stod(row[j + 3]));
    }
    return m;
}

// "section.key" -> value text, for files of the form   section:\n  key: value\n  list: [a, b, c]
std::map<std::string, std::string> readFlatYaml(const std::string& path) {
    std::ifstream in(path);
    if (!in) { std::cerr << "cannot open " << path << std::endl; std::exit(2); }
    std::map<std::string, std::string> kv;
    std::string line, section;
    while (std::getline(in, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        const size_t first = line.find_first_not_of(" \t\r");
        if (first == std::string::npos) continue;
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        std::string key = line.substr(first, colon - first), val = line.substr(colon + 1);
        const size_t a = val.find_first_not_of(" \t\r"), b = val.find_last_not_of(" \t\r");
        val = a == std::string::npos ? "" : val.substr(a, b - a + 1);
        if (first == 0) { section = key; continue; }
        kv[section + "." + key] = val;
    }
    return kv;
}

std::vector<std::string> listOf(std::string v) {
    for (char& c : v) if (c == '[' || c == ']' || c == '"' || c == '\'') c = ' ';
    std::vector<std::string> out;
    std::stringstream ss(v);
    std::string cell;
    while (std::getline(ss, cell, ',')) {
        const size_t a = cell.find_first_not_of(" \t"), b = cell.find_last_not_of(" \t");
        if (a != std::string::npos) out.push_back(cell.substr(a, b - a + 1));
    }
    return out;
}

// The keys VIOFilter::Settings(const YAML::Node&) reads (VIOFilterSettings.h:56-109); absent keys keep the defaults.
VIOFilter::Settings settingsOf(const std::map<std::string, std::string>& kv) {
    VIOFilter::Settings s;
    auto num = [&](const char* k, double& d) { auto it = kv.find(std::string("eqf.") + k); if (it != kv.end()) d = std::stod(it->second); };
    auto flag = [&](const char* k, bool& d) { auto it = kv.find(std::string("eqf.") + k); if (it != kv.end()) d = (it->second == "true" || it->second == "True" || it->second == "1"); };
    num("biasOmegaProcessVariance", s.biasOmegaProcessVariance); num("biasAccelProcessVariance", s.biasAccelProcessVariance);
    num("gravityProcessVariance", s.gravityProcessVariance); num("velocityProcessVariance", s.velocityProcessVariance);
    num("pointProcessVariance", s.pointProcessVariance); num("measurementVariance", s.measurementVariance);
    num("velOmegaVariance", s.velOmegaVariance); num("velAccelVariance", s.velAccelVariance);
    num("initialGravityVariance", s.initialGravityVariance); num("initialVelocityVariance", s.initialVelocityVariance);
    num("initialPointVariance", s.initialPointVariance); num("initialBiasOmegaVariance", s.initialBiasOmegaVariance);
    num("initialBiasAccelVariance", s.initialBiasAccelVariance);
    flag("useInnovationLift", s.useInnovationLift); flag("useDiscreteInnovationLift", s.useDiscreteInnovationLift);
    flag("useDiscreteVelocityLift", s.useDiscreteVelocityLift); flag("fastRiccati", s.fastRiccati);
    num("outlierThreshold", s.outlierThreshold); num("initialSceneDepth", s.initialSceneDepth);
    auto vec3 = [&](const char* k, Eigen::Vector3d& d) {
        auto it = kv.find(std::string("eqf.") + k);
        if (it == kv.end()) return;
        const std::vector<std::string> l = listOf(it->second);
        d = Eigen::Vector3d(std::stod(l.at(0)), std::stod(l.at(1)), std::stod(l.at(2)));
    };
    vec3("initialAccelBias", s.initialAccelBias);
    vec3("initialOmegaBias", s.initialOmegaBias);
    auto it = kv.find("eqf.cameraOffset");
    if (it != kv.end()) {
        const std::vector<std::string> l = listOf(it->second);   // ["xw", x, y, z, qw, qx, qy, qz]
        if (l.at(0) != "xw") { std::cerr << "cameraOffset must start with xw" << std::endl; std::exit(2); }
        s.cameraOffset.x() = Eigen::Vector3d(std::stod(l.at(1)), std::stod(l.at(2)), std::stod(l.at(3)));
        s.cameraOffset.R().fromQuaternion(Eigen::Quaterniond(std::stod(l.at(4)), std::stod(l.at(5)), std::stod(l.at(6)), std::stod(l.at(7))));
    }
    return s;
}

template <class Filter>
int replay(Filter& filter, const std::vector<Row>& imuRows, const std::vector<Row>& measRows, double startTime, bool writeState,
           bool writeFilter, const std::string& outState, const std::string& outFilter, int precision) {
    std::ofstream outputFile, internalFile;
    if (writeState) { outputFile.open(outState); outputFile << "time, tx, ty, tz, qw, qx, qy, qz, vx, vy, vz, N, p1id, p1x, p1y, p1z, ..." << std::endl; }
    if (writeFilter) { internalFile.open(outFilter); internalFile << "time, t0x, ..., Sigma(5+3N, 5+3N)" << std::endl; }
    if (imuRows.empty() || measRows.empty()) return 1;
    size_t imuIter = 0, measIter = 0;
    IMUVelocity imuData = imuOf(imuRows[0]);
    VisionMeasurement measData = measOf(measRows[0]);
    int imuDataCounter = 0, visionDataCounter = 0;
    while (true) {
        if (imuData.stamp < measData.stamp) {
            if (imuData.stamp > startTime) {
                filter.processIMUData(imuData);
                ++imuDataCounter;
            }
            if (++imuIter == imuRows.size()) break;
            imuData = imuOf(imuRows[imuIter]);
        } else {
            if (measData.stamp > startTime) {
                filter.processVisionData(measData);
                ++visionDataCounter;
            }
            VIOState estimatedState = filter.stateEstimate();
            if (writeState)
                outputFile << std::setprecision(20) << filter.getTime() << std::setprecision(precision) << ", " << estimatedState << std::endl;
            if (writeFilter)
                internalFile << std::setprecision(20) << filter.getTime() << std::setprecision(precision) << ", " << filter << std::endl;
            if (++measIter == measRows.size()) break;
            measData = measOf(measRows[measIter]);
        }
    }
    std::cout << "Processed " << imuDataCounter << " IMU and " << visionDataCounter << " vision measurements." << std::endl;
    return 0;
}

template <class Filter>
int run(const VIOFilter::Settings& settings, bool aux, const std::vector<Row>& imuRows, const std::vector<Row>& measRows, double startTime,
        bool ws, bool wf, const std::string& os, const std::string& of, int precision) {
    if (!aux) {
        Filter filter(settings);                       // main.cpp:85-86
        return replay(filter, imuRows, measRows, startTime, ws, wf, os, of, precision);
    }
    // the other constructors and set-up calls of VIOFilter.h:69-79
    AuxiliaryFilterData a;
    a.initialAttitude = Eigen::Quaterniond(0.96, 0.2, -0.14, 0.12);
    const double n = a.initialAttitude.norm();
    a.initialAttitude = Eigen::Quaterniond(a.initialAttitude.w() / n, a.initialAttitude.x() / n, a.initialAttitude.y() / n, a.initialAttitude.z() / n);
    a.initialPosition = Eigen::Vector3d(0.5, -0.25, 1.5);
    a.initialTime = 0.0;
    a.cameraOffset = settings.cameraOffset;
    Filter filter(a, settings);
    Filter moved(std::move(filter));                   // move construction, then move assignment (eqf_vio_ros_node.cpp:59)
    filter = std::move(moved);
    a.initialPosition = Eigen::Vector3d(0.0, 0.0, 0.0);
    filter.setAuxiliaryData(a);
    if (!imuRows.empty()) filter.initialiseFromIMUData(imuOf(imuRows[0]));
    filter.settings->measurementVariance *= 2.0;       // settings are read at use time through the public pointer
    const Eigen::MatrixXd S0 = filter.stateCovariance();
    if (S0.rows() != SIGMA_BASE_SIZE || S0.cols() != SIGMA_BASE_SIZE) return 3;
    return replay(filter, imuRows, measRows, startTime, ws, wf, os, of, precision);
}

}  // namespace

int main(int argc, char** argv) {
    std::string impl = "ref";
    std::vector<std::string> pos;
    int precision = 5;
    bool aux = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--impl" && i + 1 < argc) impl = argv[++i];
        else if (a == "--precision" && i + 1 < argc) precision = std::atoi(argv[++i]);
        else if (a == "--aux") aux = true;
        else pos.push_back(a);
    }
    if (pos.size() != 5) {
        std::cout << "Usage: dropin_replay --impl ref|b200 IMU_file meas_file config_file out_state out_filter [--precision P] [--aux]" << std::endl;
        return 1;
    }
    const std::vector<Row> imuRows = readRows(pos[0]), measRows = readRows(pos[1]);
    const std::map<std::string, std::string> kv = readFlatYaml(pos[2]);
    auto get = [&](const char* k, const char* dflt) { auto it = kv.find(k); return it == kv.end() ? std::string(dflt) : it->second; };
    const double startTime = std::stod(get("main.startTime", "0"));
    const bool writeState = get("main.writeState", "true") == "true", writeFilter = get("main.writeFilter", "true") == "true";
    const VIOFilter::Settings settings = settingsOf(kv);
    try {
        if (impl == "b200") return run<VIOFilterB200>(settings, aux, imuRows, measRows, startTime, writeState, writeFilter, pos[3], pos[4], precision);
        return run<VIOFilter>(settings, aux, imuRows, measRows, startTime, writeState, writeFilter, pos[3], pos[4], precision);
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what() << std::endl;
        return 4;
    }
}

// heros_host.hpp -- C++ host side above the C ABI of libheros_gpu.so (include/heros_gpu.h): the plug-in classes the
// reference registers with the medlabs engine, under their own names, with the same methods, argument meaning and error
// behaviour, backed by the GPU engine.  The reference is Java (eu.heros.disease.*, eu.heros.policy.*); a JVM host binds
// the C ABI through Panama FFM.  This header is the same layer for a compiled host without a JVM and
// what tests/cpp/test_host.cpp drives; it is header-only and contains no CUDA -- every computation happens behind the
// C ABI, and without a usable device the HerosModel constructor throws.
//
//   eu/heros/model/HerosModel.java            -> heros::HerosModel   (parameter map, simulator clock, the engine handle)
//   medlabs InfectionRecord                   -> heros::InfectionRecord
//   medlabs DiseaseTransmission (abstract)    -> heros::DiseaseTransmission
//   eu/heros/disease/Covid19TransmissionArea  -> heros::Covid19TransmissionArea      (java :59-69, :133-255)
//   eu/heros/disease/Covid19TransmissionDistance -> heros::Covid19TransmissionDistance (java :55-69, :115-264)
//   eu/heros/disease/Covid19Progression       -> heros::Covid19Progression            (java :126-201)
//   eu/heros/policy/DiseasePolicy             -> heros::DiseasePolicy                 (java :18-43)
//   eu/heros/policy/LocationPolicy            -> heros::LocationPolicy                (java :19-46)
//   factory/ConstructHerosModel.java:136-146  -> heros::constructDiseaseModel
#pragma once

#include "heros_gpu.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace heros {

// medlabs' checked MedlabsException (Covid19Progression's constructor declares it, java :126) and the unchecked
// MedlabsRuntimeException (HerosModel.java:494,501)
struct MedlabsException : std::runtime_error { using std::runtime_error::runtime_error; };
struct MedlabsRuntimeException : std::runtime_error { using std::runtime_error::runtime_error; };
// a non-zero status of the C ABI
struct EngineError : std::runtime_error
{
    int32_t code;
    EngineError(int32_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};

namespace detail {
inline std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char) s[a])) ++a;
    while (b > a && std::isspace((unsigned char) s[b - 1])) --b;
    return s.substr(a, b - a);
}
inline std::string lower(std::string s)
{
    for (char& c : s) c = (char) std::tolower((unsigned char) c);
    return s;
}
inline std::vector<std::string> split(const std::string& s, char sep)
{
    std::vector<std::string> out;
    std::string item;
    std::istringstream in(s);
    while (std::getline(in, item, sep)) out.push_back(item);
    return out;
}
inline double number(const std::string& text, const std::string& context)
{
    const std::string s = trim(text);
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (s.empty() || end != s.c_str() + s.size()) throw MedlabsException("cannot parse number '" + text + "' in " + context);
    return v;
}
}  // namespace detail

using Parameters = std::map<std::string, std::string>;

// A Java .properties file: "key = value", '#' / '!' comments (the files under src/main/resources, loaded by
// HerosApplication.java:146-150).
inline Parameters loadProperties(const std::string& path)
{
    std::ifstream in(path);
    if (!in) throw MedlabsException("cannot read " + path);
    Parameters out;
    std::string line;
    while (std::getline(in, line))
    {
        line = detail::trim(line);
        if (line.empty() || line[0] == '#' || line[0] == '!') continue;
        size_t k = 0;
        while (k < line.size() && line[k] != '=' && line[k] != ':' && !std::isspace((unsigned char) line[k])) ++k;
        std::string key = line.substr(0, k), rest = detail::trim(line.substr(k));
        if (!rest.empty() && (rest[0] == '=' || rest[0] == ':')) rest = detail::trim(rest.substr(1));
        out[key] = rest;
    }
    return out;
}

// ConditionalProbability strings of the covidP.* fractions: a number, "age{0-19: 0.8, 20-55: 0.5}" or
// "gender{M:0.45, F:0.5}" (alpha-area.properties:42-47), tabulated as table[female][age 0..127].  Ages outside every
// band get 0; with overlapping bands the first match wins (medlabs' tie rule is not visible in the reference; the
// shipped files never overlap).
inline void parseConditionalProbability(const std::string& text, double table[2][128])
{
    const std::string s = detail::trim(text);
    for (int g = 0; g < 2; ++g)
        for (int a = 0; a < 128; ++a) table[g][a] = 0.0;
    const size_t open = s.find('{');
    if (open == std::string::npos)
    {
        const double v = detail::number(s, text);
        for (int g = 0; g < 2; ++g)
            for (int a = 0; a < 128; ++a) table[g][a] = v;
        return;
    }
    const std::string kind = detail::lower(detail::trim(s.substr(0, open)));
    if (s.back() != '}' || (kind != "age" && kind != "gender")) throw MedlabsException("cannot parse probability '" + text + "'");
    const std::string body = s.substr(open + 1, s.size() - open - 2);
    bool done[128] = {};
    for (const std::string& entry : detail::split(body, ','))
    {
        if (detail::trim(entry).empty()) continue;
        const std::vector<std::string> kv = detail::split(entry, ':');
        if (kv.size() != 2) throw MedlabsException("cannot parse probability '" + text + "'");
        const double v = detail::number(kv[1], text);
        if (kind == "age")
        {
            const std::vector<std::string> band = detail::split(kv[0], '-');
            if (band.size() != 2) throw MedlabsException("cannot parse age band in '" + text + "'");
            const int lo = (int) detail::number(band[0], text), hi = (int) detail::number(band[1], text);
            for (int a = std::max(lo, 0); a <= std::min(hi, 127); ++a)
                if (!done[a]) { table[0][a] = table[1][a] = v; done[a] = true; }
        }
        else
        {
            const std::string g = detail::lower(detail::trim(kv[0]));
            if (g.empty() || (g[0] != 'm' && g[0] != 'f')) throw MedlabsException("unknown gender key in '" + text + "'");
            for (int a = 0; a < 128; ++a) table[g[0] == 'f' ? 1 : 0][a] = v;
        }
    }
}

// medlabs DistributionParser.parseDistContinuous as used by Covid19Progression.java:144-173: "Triangular(a,b,c)",
// "Uniform(a,b)", "Constant(x)", parameters in days.
inline heros_duration parseDistribution(const std::string& text)
{
    const std::string s = detail::trim(text);
    const size_t open = s.find('(');
    if (open == std::string::npos || s.back() != ')') throw MedlabsException("cannot parse distribution '" + text + "'");
    const std::string name = detail::lower(detail::trim(s.substr(0, open)));
    std::vector<double> args;
    for (const std::string& a : detail::split(s.substr(open + 1, s.size() - open - 2), ','))
        if (!detail::trim(a).empty()) args.push_back(detail::number(a, text));
    heros_duration d{};
    if ((name == "triangular" || name == "tria" || name == "triang") && args.size() == 3)
    { d.kind = HEROS_DIST_TRIANGULAR; d.a = args[0]; d.b = args[1]; d.c = args[2]; return d; }
    if ((name == "uniform" || name == "unif") && args.size() == 2)
    { d.kind = HEROS_DIST_UNIFORM; d.a = args[0]; d.b = args[1]; return d; }
    if ((name == "constant" || name == "const") && args.size() == 1)
    { d.kind = HEROS_DIST_CONSTANT; d.a = args[0]; return d; }
    throw MedlabsException("distribution '" + text + "' is not supported by the GPU engine");
}

// ---------------------------------------------------------------------------------------------------------------------
// HerosModel: the slice of MedlabsModelInterface the disease plug-ins use -- the parameter map, the simulator clock --
// and, instead of the person / location object maps, the GPU engine that holds that state.
// ---------------------------------------------------------------------------------------------------------------------
class DiseaseTransmission;
class Covid19Progression;

class HerosModel
{
public:
    explicit HerosModel(Parameters parameters, int64_t seed = -1, int device = 0, int trigger = HEROS_TRIGGER_EXIT_ONLY,
                        double windowHours = 24.0, uint32_t flags = 0)
        : parameters_(std::move(parameters))
    {
        std::string name = "area";
        if (parameters_.count("generic.diseasePropertiesModel")) name = detail::lower(detail::trim(parameters_["generic.diseasePropertiesModel"]));
        if (name != "area" && name != "distance")  // ConstructHerosModel.java:137-141
            throw MedlabsException("generic.diseasePropertiesModel must be area or distance, not " + name);
        heros_config cfg{};
        cfg.abi_version = HEROS_ABI_VERSION;
        cfg.model = modelKind_ = name == "area" ? HEROS_MODEL_AREA : HEROS_MODEL_DISTANCE;
        cfg.trigger = trigger;
        cfg.device = device;
        cfg.seed = seed >= 0 ? (uint64_t) seed
                             : (uint64_t) (parameters_.count("generic.Seed") ? detail::number(parameters_["generic.Seed"], "generic.Seed") : 111);
        cfg.window_hours = windowHours;
        cfg.flags = flags;
        const int32_t rc = heros_create(&cfg, &engine_);
        if (rc != HEROS_OK) throw EngineError(rc, std::string("heros_create: ") + heros_last_error(nullptr));  // no CPU fallback
    }
    ~HerosModel() { if (engine_) heros_destroy(engine_); }
    HerosModel(const HerosModel&) = delete;
    HerosModel& operator=(const HerosModel&) = delete;

    // MedlabsModelInterface.getParameterValue* (Covid19TransmissionArea.java:63-68)
    const std::string& getParameterValue(const std::string& key) const
    {
        auto it = parameters_.find(key);
        if (it == parameters_.end()) throw MedlabsRuntimeException("parameter " + key + " not found");
        return it->second;
    }
    double getParameterValueDouble(const std::string& key) const { return detail::number(getParameterValue(key), key); }
    int getParameterValueInt(const std::string& key) const { return (int) detail::number(getParameterValue(key), key); }
    double getSimulatorTime() const
    {
        double t = 0.0;
        check(heros_get_time(engine_, &t));
        return t;
    }
    // ConstructHerosModel.java:143-144
    void setDiseaseProgression(std::shared_ptr<Covid19Progression> p) { progression_ = std::move(p); }
    void setDiseaseTransmission(std::shared_ptr<DiseaseTransmission> t) { transmission_ = std::move(t); }
    const std::shared_ptr<DiseaseTransmission>& getDiseaseTransmission() const { return transmission_; }
    const std::shared_ptr<Covid19Progression>& getDiseaseProgression() const { return progression_; }
    // model.getLocationTypeNameMap() (LocationPolicy.java:23): type name -> id in the order of locationtypes.csv
    void setLocationTypeNames(const std::vector<std::string>& names)
    {
        typeNames_.clear();
        for (size_t i = 0; i < names.size(); ++i) typeNames_[names[i]] = (int) i;
    }
    const std::map<std::string, int>& getLocationTypeNameMap() const { return typeNames_; }

    // The locations ConstructHerosModel.readLocationTable builds as LocationProbBased (factory/ConstructHerosModel.java:382-388,
    // rows of epidemiology/infection_rates.csv) collect here and go to the engine in one call, after heros_set_locations.
    void addProbBasedLocation(int locationId, double infectionRateFactor, double infectionRate)
    {
        probLoc_.push_back(locationId); probFactor_.push_back(infectionRateFactor); probRate_.push_back(infectionRate);
    }
    void installProbBasedLocations()
    {
        check(heros_set_rate_locations(engine_, (int32_t) probLoc_.size(), probLoc_.data(), probFactor_.data(), probRate_.data()));
    }

    heros_handle engine() const { return engine_; }
    int modelKind() const { return modelKind_; }
    void check(int32_t rc) const
    {
        if (rc != HEROS_OK) throw EngineError(rc, heros_last_error(engine_));
    }

private:
    Parameters parameters_;
    heros_handle engine_ = nullptr;
    int modelKind_ = HEROS_MODEL_AREA;
    std::shared_ptr<DiseaseTransmission> transmission_;
    std::shared_ptr<Covid19Progression> progression_;
    std::map<std::string, int> typeNames_;
    std::vector<int32_t> probLoc_;
    std::vector<double> probFactor_, probRate_;
};

// medlabs LocationProbBased(model, locationId, locationType, lat, lon, nbSublocations, area, infectionRateFactor, infectionRate,
// referenceGroupMap, Covid19Progression.exposed) as the factory constructs it (ConstructHerosModel.java:386-387): of its
// arguments the engine needs the id and the two rates (the reference group is the Worker role, :303-308).
class LocationProbBased
{
public:
    LocationProbBased(HerosModel& model, int locationId, double infectionRateFactor, double infectionRate)
    {
        model.addProbBasedLocation(locationId, infectionRateFactor, infectionRate);
    }
};

// medlabs InfectionRecord(exposurePhase, location) as infectPeople fills it (Covid19TransmissionArea.java:135-193)
struct InfectionRecord
{
    std::string exposurePhase;
    int location = -1;
    bool calculated = false;
    std::vector<int32_t> infectiousPersons, infectedPersons;
    double pInfection = std::nan("");
    bool isCalculated() const { return calculated; }
};

// medlabs abstract DiseaseTransmission(model, name)
class DiseaseTransmission
{
public:
    DiseaseTransmission(HerosModel& model, std::string name) : model_(model), name_(std::move(name)) {}
    virtual ~DiseaseTransmission() = default;
    const std::string& getName() const { return name_; }

    // infectPeople(Location, TIntSet personsInSublocation, double duration): one evaluation, on the GPU, against the
    // engine's agent state.  allPersonIds = location.getAllPersonIds() (read by the total-location branch only);
    // now = model.getSimulator().getSimulatorTime().  Never returns "null": a fresh record, as in the reference.
    InfectionRecord infectPeople(int location, const std::vector<int32_t>& personsInSublocation, double duration,
                                 const std::vector<int32_t>* allPersonIds = nullptr, double now = std::nan("")) const
    {
        InfectionRecord rec;
        rec.exposurePhase = "Exposed";
        rec.location = location;
        const double t = now == now ? now : model_.getSimulatorTime();
        const size_t cap = std::max(personsInSublocation.size(), allPersonIds ? allPersonIds->size() : (size_t) 0) + 1;
        std::vector<int32_t> infectious(cap), infected(cap);
        int32_t calculated = 0, nInfectious = 0, nInfected = 0;
        model_.check(heros_infect_people(model_.engine(), location, (int32_t) personsInSublocation.size(),
                                         personsInSublocation.data(), allPersonIds ? (int32_t) allPersonIds->size() : 0,
                                         allPersonIds ? allPersonIds->data() : nullptr, t, duration, &calculated,
                                         infectious.data(), &nInfectious, infected.data(), &nInfected, &rec.pInfection));
        rec.calculated = calculated != 0;
        rec.infectiousPersons.assign(infectious.begin(), infectious.begin() + nInfectious);
        rec.infectedPersons.assign(infected.begin(), infected.begin() + nInfected);
        return rec;
    }

    // setParameter(String, double): unknown names give a message on stderr, no exception
    // (Covid19TransmissionArea.java:254, Covid19TransmissionDistance.java:263)
    void setParameter(const std::string& parameterName, double value, double atTime = std::nan(""))
    {
        const double t = atTime == atTime ? atTime : model_.getSimulatorTime();
        const int32_t rc = heros_set_parameter(model_.engine(), parameterName.c_str(), value, t);
        if (rc == HEROS_ERR_UNKNOWN_PARAMETER) std::cerr << "Unrecognized variable name " << parameterName << std::endl;
        else model_.check(rc);
    }

protected:
    HerosModel& model_;
    std::string name_;
};

// Covid19TransmissionArea(model): reads covidT_area.* (java :63-68; the engine converts days -> hours, seconds -> hours)
class Covid19TransmissionArea : public DiseaseTransmission
{
public:
    explicit Covid19TransmissionArea(HerosModel& model) : DiseaseTransmission(model, "Covid19")
    {
        params.contagiousness = model.getParameterValueDouble("covidT_area.contagiousness");
        params.beta = model.getParameterValueDouble("covidT_area.beta");
        params.t_e_min = model.getParameterValueDouble("covidT_area.t_e_min");
        params.t_e_mode = model.getParameterValueDouble("covidT_area.t_e_mode");
        params.t_e_max = model.getParameterValueDouble("covidT_area.t_e_max");
        params.calculation_threshold = model.getParameterValueDouble("covidT_area.calculation_threshold");
        model.check(heros_set_transmission_area(model.engine(), &params));
    }
    heros_area_params params{};
};

// Covid19TransmissionDistance(model): reads covidT_dist.* (java :59-68); psi and mu change at run time (java :251-264)
class Covid19TransmissionDistance : public DiseaseTransmission
{
public:
    explicit Covid19TransmissionDistance(HerosModel& model) : DiseaseTransmission(model, "Covid19")
    {
        params.L = model.getParameterValueDouble("covidT_dist.L");
        params.I = model.getParameterValueDouble("covidT_dist.I");
        params.C = model.getParameterValueDouble("covidT_dist.C");
        params.v_max = model.getParameterValueDouble("covidT_dist.v_max");
        params.v_0 = model.getParameterValueDouble("covidT_dist.v_0");
        params.r = model.getParameterValueDouble("covidT_dist.r");
        params.psi = model.getParameterValueDouble("covidT_dist.psi");
        params.alpha = model.getParameterValueDouble("covidT_dist.alpha");
        params.mu = model.getParameterValueDouble("covidT_dist.mu");
        params.calculation_threshold = model.getParameterValueDouble("covidT_dist.calculation_threshold");
        model.check(heros_set_transmission_distance(model.engine(), &params));
    }
    heros_dist_params params{};
};

// Covid19Progression(model) throws MedlabsException: the 8 phases in registration order (java :130-137), the five
// fractions and ten duration distributions of covidP.* (java :143-173)
class Covid19Progression
{
public:
    static constexpr const char* susceptible = "Susceptible";
    static constexpr const char* exposed = "Exposed";
    static constexpr const char* infected_asymptomatic = "Infected-Asymptomatic";
    static constexpr const char* infected_symptomatic = "Infected-Symptomatic";
    static constexpr const char* hospitalized = "Hospitalized";
    static constexpr const char* icu = "ICU";
    static constexpr const char* dead = "Dead";
    static constexpr const char* recovered = "Recovered";

    explicit Covid19Progression(HerosModel& model) : model_(model), params_(new heros_prog_params())
    {
        static const char* fractionKeys[HEROS_N_FRACTIONS] = {
            "covidP.FractionAsymptomatic", "covidP.FractionSymptomaticToHospitalized", "covidP.FractionHospitalizedToICU",
            "covidP.FractionHospitalizedToDead", "covidP.FractionICUToDead"};
        static const char* periodKeys[HEROS_N_PERIODS] = {
            "covidP.IncubationPeriodAsymptomatic", "covidP.IncubationPeriodSymptomatic", "covidP.PeriodAsymptomaticToRecovered",
            "covidP.PeriodSymptomaticToHospitalized", "covidP.PeriodSymptomaticToRecovered", "covidP.PeriodHospitalizedToICU",
            "covidP.PeriodHospitalizedToDead", "covidP.PeriodHospitalizedToRecovered", "covidP.PeriodICUToDead",
            "covidP.PeriodICUToRecovered"};
        try
        {
            for (int f = 0; f < HEROS_N_FRACTIONS; ++f)
                parseConditionalProbability(model.getParameterValue(fractionKeys[f]), params_->fraction[f]);
            for (int p = 0; p < HEROS_N_PERIODS; ++p) params_->period[p] = parseDistribution(model.getParameterValue(periodKeys[p]));
        }
        catch (const MedlabsRuntimeException& e) { throw MedlabsException(e.what()); }
        model.check(heros_set_progression(model.engine(), params_.get()));
    }
    const std::string& getName() const { return name_; }
    std::vector<std::string> getDiseasePhases() const
    {
        return {susceptible, exposed, infected_asymptomatic, infected_symptomatic, hospitalized, icu, dead, recovered};
    }
    // expose(Person, DiseasePhase) (java :182-201): only Susceptible persons change; exposure time = (float) now
    void expose(const std::vector<int32_t>& persons, const std::string& exposurePhase = exposed, double atTime = std::nan(""))
    {
        if (exposurePhase != exposed) throw MedlabsRuntimeException("expose() only takes the Exposed phase");
        const double t = atTime == atTime ? atTime : model_.getSimulatorTime();
        const std::vector<double> at(persons.size(), t);
        model_.check(heros_expose(model_.engine(), (int32_t) persons.size(), persons.data(), at.data()));
    }
    void expose(int32_t person, const std::string& exposurePhase = exposed, double atTime = std::nan(""))
    {
        expose(std::vector<int32_t>{person}, exposurePhase, atTime);
    }
    const heros_prog_params& params() const { return *params_; }

private:
    HerosModel& model_;
    std::string name_ = "Covid19";
    std::unique_ptr<heros_prog_params> params_;
};

// ConstructHerosModel.java:136-146: progression first, then the transmission model selected by
// generic.diseasePropertiesModel, both registered on the model
inline void constructDiseaseModel(HerosModel& model)
{
    auto progression = std::make_shared<Covid19Progression>(model);
    std::shared_ptr<DiseaseTransmission> transmission;
    if (model.modelKind() == HEROS_MODEL_AREA) transmission = std::make_shared<Covid19TransmissionArea>(model);
    else transmission = std::make_shared<Covid19TransmissionDistance>(model);
    model.setDiseaseProgression(progression);
    model.setDiseaseTransmission(transmission);
}

// DiseasePolicy(model, time, parameterName, value) (java :18-43).  The reference schedules changeParameter on the DSOL
// event list for absolute time `time`; the engine keeps a time-stamped parameter schedule, so the change is registered
// at once and applies to every evaluation at simulator time >= time.
class DiseasePolicy
{
public:
    DiseasePolicy(HerosModel& model, double time, const std::string& parameterName, double value) : model_(model)
    {
        if (parameterName != "covidT_dist.psi" && parameterName != "covidT_dist.mu")
        {
            std::cerr << "Unrecognized variable name " << parameterName << std::endl;  // java :21-25
            return;
        }
        changeParameter(parameterName, value, time);
    }
    void changeParameter(const std::string& parameterName, double value, double time)
    {
        if (parameterName == "covidT_dist.psi") model_.getDiseaseTransmission()->setParameter("psi", value, time);
        else if (parameterName == "covidT_dist.mu") model_.getDiseaseTransmission()->setParameter("mu", value, time);
        else std::cerr << "Unrecognized variable name " << parameterName << std::endl;
    }

private:
    HerosModel& model_;
};

// LocationPolicy(model, time, locationTypeName, fractionOpen, fractionActivities, alternativeLocationName,
// reportAsLocationName) (java :19-46).  Unknown type names: message on stderr, policy ignored (java :24-34).  The event
// the reference schedules for `time` is fired by the host's day loop through openCloseLocationType(), which hands the
// closure to heros_set_closure_policy (= LocationType.setClosurePolicy, java :45) for the device-side schedule.
class LocationPolicy
{
public:
    LocationPolicy(HerosModel& model, double time, const std::string& locationTypeName, double fractionOpen,
                   double fractionActivities, const std::string& alternativeLocationName, std::string reportAsLocationName)
        : model_(model), time(time), fractionOpen(fractionOpen), fractionActivities(fractionActivities),
          reportAsLocationName(std::move(reportAsLocationName))
    {
        const auto& names = model.getLocationTypeNameMap();
        auto lt = names.find(locationTypeName);
        if (lt == names.end())
        {
            std::cerr << "LocationPolicy: Did not recognize locationType " << locationTypeName << std::endl;
            return;
        }
        auto alt = names.find(alternativeLocationName);
        if (alt == names.end())
        {
            std::cerr << "LocationPolicy: Did not recognize alternativeLocationType " << alternativeLocationName << std::endl;
            return;
        }
        locationType = lt->second;
        alternativeLocationType = alt->second;
        scheduled = true;
    }
    void openCloseLocationType() const
    {
        if (!scheduled) return;
        model_.check(heros_set_closure_policy(model_.engine(), locationType, fractionOpen, fractionActivities, alternativeLocationType));
    }

private:
    HerosModel& model_;

public:
    double time;
    int locationType = -1, alternativeLocationType = -1;
    double fractionOpen, fractionActivities;
    std::string reportAsLocationName;
    bool scheduled = false;
};

}  // namespace heros
